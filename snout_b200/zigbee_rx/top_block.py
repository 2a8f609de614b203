#!/usr/bin/env python3
# -*- coding: utf-8 -*-
"""Copy (or symlink) this file to ~/.scapy-radio/Zigbee/<hardware>/Zigbee_rx/top_block.py.

scapy-radio starts every `.py` modulation as `python2 <file> [params]`
(scapy/modules/gnuradio.py:487-489).  This file is therefore valid Python 2 AND 3: under
Python 2 it re-executes itself with python3, under Python 3 it runs the B200 receive engine
behind the same XMLRPC / UDP interface as the reference flowgraph.  The IQ source comes from
the environment (SNOUT_B200_IQ, SNOUT_B200_IQ_FORMAT, SNOUT_B200_WIDEBAND) or from --iq.
Set SNOUT_B200_ROOT to the checkout if this file was copied rather than symlinked."""
import os
import sys

if sys.version_info[0] < 3:
    os.execvp("python3", ["python3", os.path.abspath(__file__)] + sys.argv[1:])

_root = os.environ.get("SNOUT_B200_ROOT") or os.path.dirname(os.path.dirname(os.path.dirname(os.path.realpath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
from snout_b200.zigbee_rx.flowgraph import main, top_block  # noqa: E402,F401

if __name__ == "__main__":
    sys.exit(main())
