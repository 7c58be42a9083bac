from smokephysai_b200.navier_stokes import NavierStokesSimulator  # noqa: F401
