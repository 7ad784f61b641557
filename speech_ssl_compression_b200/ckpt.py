"""Checkpoint writing under data parallelism: every rank holds identical parameters and prune decisions, so only
rank 0 of a multi-process run writes files (the reference is single-process, runner.py:329-356 / 448-458)."""
import os

import torch


def is_writer():
    return int(os.environ.get("RANK", "0")) == 0


def save(states, path):
    """``torch.save`` on rank 0, a no-op elsewhere.  Returns True when the file was written."""
    if not is_writer():
        return False
    torch.save(states, path)
    return True
