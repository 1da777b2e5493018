"""Empty stub: the reference's utils.py imports pyplot/colors at module scope; plotting is out of scope."""
