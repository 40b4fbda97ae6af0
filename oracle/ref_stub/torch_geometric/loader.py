class DataLoader:  # imported by the reference's utils.py, never used on the hot path
    def __init__(self, *a, **k):
        raise NotImplementedError("stub")
