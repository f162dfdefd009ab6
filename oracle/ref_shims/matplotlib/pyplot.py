def __getattr__(name):
    raise RuntimeError("matplotlib is not available in this image (shim): " + name)
