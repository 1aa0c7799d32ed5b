"""Instance generators moved to the neutral `instances` package (inputs are shared by the oracle, the tests and
bench.py); this module keeps the oracle's historical import path."""
from instances.generators import *  # noqa: F401,F403
from instances.generators import _index_map  # noqa: F401
