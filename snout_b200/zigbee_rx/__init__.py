"""Drop-in for snout/modulations/Zigbee/<hw>/Zigbee_rx (see flowgraph.py)."""
from .flowgraph import top_block, main  # noqa: F401
