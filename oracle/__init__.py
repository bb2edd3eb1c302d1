"""CPU oracle -- test infrastructure only (see yolo_head_oracle.py header)."""
