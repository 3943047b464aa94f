import sys

from . import run

if len(sys.argv) < 3:
    raise SystemExit("usage: python -m gomavatar_b200.compat /path/to/GoMAvatar <train.py|eval.py|train_pose.py> [script args]")
run(sys.argv[1], sys.argv[2], sys.argv[3:])
