def __getattr__(name):  # any plt.<fn> is a no-op
    def _noop(*a, **k):
        return None
    return _noop
