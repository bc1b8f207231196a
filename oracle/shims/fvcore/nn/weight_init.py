def c2_msra_fill(module):  # import-only (models_painter.py:17)
    raise NotImplementedError
