class _Config(object):
    def update(self, *a, **k):
        pass


config = _Config()
