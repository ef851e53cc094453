"""Drop-in for the reference's training entry point (reference scripts/train.py:19-35):

    python scripts/train.py --config_path input_configs/train_synthetic.yaml --optim.max_train_steps 10 --log.overwrite_ok

Same flow: fix the seeds, prepare the experiment directories, `Coach(cfg).train()`.  pyrallis' `@wrap()` argument syntax
(yaml file + dotted overrides) is parsed by view_neti_b200.training.config.parse_args.
"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from view_neti_b200.training.coach import Coach
from view_neti_b200.training.config import RunConfig, parse_args


def fixseed(seed: int) -> None:
    """reference utils/fixseed.py:5-12"""
    torch.backends.cudnn.benchmark = False
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def prepare_directories(cfg: RunConfig) -> None:
    cfg.log.exp_dir = cfg.log.exp_dir / cfg.log.exp_name
    if os.path.exists(cfg.log.exp_dir) and not cfg.log.overwrite_ok:
        raise ValueError(f"Experiment folder already exists and overwrite_ok=False: [{cfg.log.exp_dir}]"
                         f" to overwrite the old experiment, add --log.overwrite_ok")
    cfg.log.exp_dir.mkdir(parents=True, exist_ok=True)
    cfg.log.logging_dir = cfg.log.exp_dir / cfg.log.logging_dir
    cfg.log.logging_dir.mkdir(parents=True, exist_ok=True)


def main(cfg: RunConfig):
    fixseed(cfg.seed)
    prepare_directories(cfg=cfg)
    coach = Coach(cfg)
    losses = coach.train()
    print(f"trained {coach.global_step} optimiser steps ({coach.micro_step} micro-steps); last loss {float(losses[-1]):.5f}")


if __name__ == "__main__":
    main(parse_args(sys.argv[1:]))
