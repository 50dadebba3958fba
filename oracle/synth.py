"""Re-export of the synthetic input generator (tools/synth.py) for the tests and tools that historically imported it
from here.  The generator manufactures inputs only; it is not an oracle and the product bench imports it from tools/."""
from tools.synth import *  # noqa: F401,F403
from tools.synth import CALIB_752x480, scaled_calibration, synth_pair, unrectify, synth_raw_pair  # noqa: F401
