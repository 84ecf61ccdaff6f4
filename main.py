"""Launcher: ``python main.py --experiments configs/basis_exp/experiment_fedstil.yaml [more.yaml ...]``.

Same CLI as the reference (``main.py:7-25``). Multi-GPU: ``torchrun --nproc-per-node N main.py --experiments ...``
(one rank per GPU, clients round-robined over ranks). ``--common`` selects another ``common.yaml``.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from flpr_b200.runtime.config import load_experiments  # noqa: E402
from flpr_b200.runtime.experiment import ExperimentStage  # noqa: E402


def main(argv=None) -> None:
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--experiments", type=str, nargs="+", required=True, help="Experiment yaml file path")
    parser.add_argument("--common", type=str, default="./configs/common.yaml", help="Common yaml file path")
    parser.add_argument("--synthetic", action="store_true",
                        help="use generated person crops instead of datasets_dir (no dataset is available offline)")
    args = parser.parse_args(argv)
    common, exps = load_experiments(args.common, args.experiments)
    factory = None
    if args.synthetic:
        from flpr_b200.data.synthetic import synthetic_source_factory
        size = tuple(common.get("defaults", {}).get("task_opts", {}).get("augment_opts", {}).get("img_size", (128, 64)))
        factory = synthetic_source_factory(size=size)
    with ExperimentStage(common, exps, source_factory=factory) as stage:
        stage.run()


if __name__ == "__main__":
    main()
