"""Drop-in import path of the reference (`from src.physics.smoke_simulator import SmokeSimulator`,
inference.py:13, data_loader.py:39).  Only src.physics is provided: the rest of the reference's `src`
package (models, utils, evaluation) is out of scope and stays the reference's own stock-PyTorch code."""
