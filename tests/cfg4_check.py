"""Runs odcp_b200.train_step's parity check (sharded training step == whole-batch step) as a script, alone or
under torchrun; prints one JSON line on rank 0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import odcp_b200  # noqa: F401  (registers the package alias)
from odcp_b200 import train_step

if __name__ == "__main__":
    sys.argv = [sys.argv[0], "--check"]
    train_step.main()
