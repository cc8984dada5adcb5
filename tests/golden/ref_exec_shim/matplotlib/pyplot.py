def __getattr__(name):
    raise RuntimeError("matplotlib stub: plotting is out of scope")
