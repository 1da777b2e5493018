class _Config:
    def update(self, *a, **k):
        pass


config = _Config()
