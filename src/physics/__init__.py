from .smoke_simulator import SmokeSimulator  # noqa: F401
from .navier_stokes import NavierStokesSimulator  # noqa: F401
from .fractal_generator import FractalGenerator  # noqa: F401
