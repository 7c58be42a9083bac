from smokephysai_b200.fractal_generator import FractalGenerator  # noqa: F401
