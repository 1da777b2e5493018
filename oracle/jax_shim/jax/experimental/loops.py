class Scope(object):
    """`with loops.Scope() as s: for n in s.range(N): s.x = ...` as a plain python loop."""
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    @staticmethod
    def range(*args):
        return range(*[int(a) for a in args])
