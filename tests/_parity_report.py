"""Measured parity numbers of the GPU tests: printed, and appended as JSON lines to $RGBD_PARITY_REPORT when set
(the committed copies live under profiles/)."""
import json
import os


def report(name: str, **values) -> None:
    path = os.environ.get("RGBD_PARITY_REPORT")
    line = json.dumps({"test": name, **values})
    print("[parity]", line)
    if path:
        with open(path, "a") as f:
            f.write(line + "\n")
