// hand-written fixture for the jsonnet-subset evaluator (not a reference file)
{
    opt: { rate: 5e-1, decay: 1e-4, flags: [true, false, null], },
    norm:: { mean: [1, 2, 3] },
    mean: self.norm.mean,
}
