"""Neutralised `typeguard` (typeguard 4.x breaks the reference's
@typechecked annotations, which collide with the local variable `B`)."""


def typechecked(f=None, **kw):
    if f is None:
        return lambda g: g
    return f
