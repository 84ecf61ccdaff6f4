"""World-size-2 (gloo) run of this engine for the golden comparison: one client per rank, the server role replicated.
    torchrun ... dist_golden_check.py <method> <dir> <rounds>
``<dir>/ref_out.pt`` holds the reference's output (initial weights); every rank writes ``<dir>/rank{r}_out.pt`` with the
checkpoint files *it* produced and its experiment log."""
import os
import pathlib
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import test_golden_experiment as T  # noqa: E402


def main():
    method, root, rounds = sys.argv[1], pathlib.Path(sys.argv[2]), int(sys.argv[3])
    rank = int(os.environ.get("RANK", 0))
    ref = torch.load(root / "ref_out.pt", weights_only=False)
    work = root / f"rank{rank}"
    work.mkdir(parents=True, exist_ok=True)
    files, log = T._run_ours(work, method, T._splits(), ref["init"], rounds)
    torch.save({"files": files, "log": log}, root / f"rank{rank}_out.pt")
    print(f"DIST_GOLDEN rank{rank} done", flush=True)


if __name__ == "__main__":
    main()
