"""TEST INFRASTRUCTURE ONLY -- `dgl.function` stand-in: only the two built-ins the reference uses
(layers.py:229-232: `fn.copy_u('h','m')`, `fn.sum('m','h')`)."""


class CopyU:
    def __init__(self, u, out):
        self.u, self.out = u, out


class Sum:
    def __init__(self, msg, out):
        self.msg, self.out = msg, out


def copy_u(u, out):
    return CopyU(u, out)


copy_src = copy_u


def sum(msg, out):  # noqa: A001  (mirrors dgl.function.sum)
    return Sum(msg, out)
