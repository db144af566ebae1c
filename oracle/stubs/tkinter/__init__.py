from _noop import module_getattr as __getattr__  # noqa: F401  (PEP 562 module-level __getattr__)
