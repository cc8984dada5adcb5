def Fire(*a, **k):
    raise RuntimeError("fire stub")
