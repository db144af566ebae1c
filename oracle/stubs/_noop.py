"""Do-nothing stand-ins used ONLY so the read-only reference modules can be imported
in a container without matplotlib / tkinter (test infrastructure, never shipped)."""


class _Noop:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Noop()

    def __getattr__(self, name):
        return _Noop()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


def module_getattr(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Noop()
