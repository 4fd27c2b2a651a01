"""Clean-room stand-in for the third-party `gym_minigrid` package (classic
1.0.x API).  TEST INFRASTRUCTURE ONLY -- see minigrid.py."""
from . import minigrid, roomgrid, wrappers  # noqa: F401
