from smokephysai_b200.smoke_simulator import SmokeSimulator  # noqa: F401
