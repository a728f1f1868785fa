local base = import "base.jsonnet";

base {
    name: "derived",  // late-bound: child.name must follow
    opt+: { rate: 0.25 },
    window+: { _unit: 3 },
}
