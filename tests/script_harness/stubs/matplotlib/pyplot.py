"""matplotlib.pyplot stand-in: any attribute is a callable that accepts everything and returns an object on which any
further attribute / call / iteration / indexing works (fig, ax = plt.subplots(); bars = plt.bar(); bar.get_height() ...)."""


class _Anything:
    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __getitem__(self, i):
        return _Anything()

    def __float__(self):
        return 0.0

    def __add__(self, o):
        return 0.0

    __radd__ = __add__


def subplots(*a, **k):
    return _Anything(), _Anything()


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Anything()
