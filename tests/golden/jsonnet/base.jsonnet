local lib = import "lib.libsonnet";
/* block comment */
{
    name: 'base',
    child: { name: $.name, twice: self.name + self.name },
    opt: lib.opt,
    mean: lib.mean,
    count: '12',
    speeds: [4],
    window: {
        _unit:: 8,
        size: if std.length($.speeds) == 0 then self._unit else $.speeds[0] * self._unit,
    },
}
