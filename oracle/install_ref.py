"""TEST / BASELINE INFRASTRUCTURE ONLY — puts the UNMODIFIED reference files of the pretraining hot path into the
git-ignored ``baseline/_ref/`` so that they travel to the GPU box with gpurun (the box has no /root/reference).

The reference is plain Python without package metadata (no setup.py / pyproject), so ``pip install --target`` has
nothing to install; this recipe is the equivalent: a byte-for-byte copy of the files SURVEY.md §8(a) cites, made from
where they lie under /root/reference, with a manifest of their sha256.  Nothing is copied into tracked paths.

    python oracle/install_ref.py          # no-op when /root/reference is absent (e.g. on the GPU box)
"""
import hashlib
import json
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference")
DST = ROOT / "baseline" / "_ref"
FILES = [
    "models/resnet.py", "models/c3d.py", "models/s3dg.py", "models/r2plus1d_vcop.py",
    "moco/split_wrapper.py", "moco/builder_diffspeed_diffloss.py",
]


def install() -> bool:
    if not (SRC / FILES[-1]).exists():
        return DST.exists()
    manifest = {}
    for rel in FILES:
        out = DST / rel
        out.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(SRC / rel, out)
        manifest[rel] = hashlib.sha256(out.read_bytes()).hexdigest()
    # the jsonnet configs of the path (loaded unchanged by rspnet_b200.config.get_config)
    for sub in ("pretrain", "model", "dataset", "optimizer"):
        for f in sorted((SRC / "config" / sub).glob("*sonnet")):
            out = DST / "config" / sub / f.name
            out.parent.mkdir(parents=True, exist_ok=True)
            shutil.copyfile(f, out)
            manifest[f"config/{sub}/{f.name}"] = hashlib.sha256(out.read_bytes()).hexdigest()
    (DST / "MANIFEST.json").write_text(json.dumps({"source": str(SRC), "sha256": manifest}, indent=1))
    return True


if __name__ == "__main__":
    ok = install()
    print(f"baseline/_ref {'ready' if ok else 'unavailable (no /root/reference and no previous install)'}")
    sys.exit(0)
