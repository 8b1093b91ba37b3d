__thr_shim__ = True
