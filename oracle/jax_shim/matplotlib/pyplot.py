def __getattr__(name):
    raise NotImplementedError("matplotlib stub: plotting is out of scope")
