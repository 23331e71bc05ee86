"""TEST INFRASTRUCTURE: import stub (viewer only; never called by the oracle harness)."""


class _Noop(object):
    def __getattr__(self, name):
        return lambda *a, **k: None


mlab = _Noop()
