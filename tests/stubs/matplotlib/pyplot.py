class _Axis:
    def imshow(self, *a, **k):
        pass

    def set_title(self, *a, **k):
        pass

    def axis(self, *a, **k):
        pass


def subplots(rows=1, cols=1, **kwargs):
    return None, [_Axis() for _ in range(rows * cols)]


def tight_layout():
    pass


def savefig(path, *a, **k):
    with open(path, "wb") as f:
        f.write(b"stub-png")


def close(*a, **k):
    pass
