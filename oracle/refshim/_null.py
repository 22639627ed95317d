"""Null object standing in for matplotlib / mpl_toolkits when an unmodified reference driver is run for
its numbers (TEST INFRASTRUCTURE ONLY): every attribute, call, item and two-way unpacking yields the same
inert object, so the plotting tail of a driver executes without drawing anything."""


class Null:
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return self

    def __call__(self, *a, **k):
        return self

    def __getitem__(self, k):
        return self

    def __iter__(self):
        return iter((self, self))


NULL = Null()
