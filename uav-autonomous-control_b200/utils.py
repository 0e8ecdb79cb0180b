"""Scheduling and flight configuration of the batched path.

Same entry points and keys as the reference's ``uav_ac/utils.py`` (``get_config`` -> the ``[DEFAULT]`` and
``[SIM_FLIGHT]`` sections of ``config.ini``: ``frequency``, ``velocity``, ``min_dist_target``; ``parse_array`` for
list-valued entries), plus the ``[BATCH]`` section that only the batched path knows (Monte-Carlo seed and ranges).
"""
from __future__ import annotations

import configparser
import functools
import json
import os

import numpy as np

CONFIG_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config.ini")


@functools.lru_cache(maxsize=None)
def _sections(path: str = CONFIG_FILE) -> configparser.ConfigParser:
    parser = configparser.ConfigParser(inline_comment_prefixes="#")
    if not parser.read(path):
        raise FileNotFoundError(path)
    return parser


def get_config():
    """(default section, flight section), exactly what ``uav_ac.utils.get_config()`` returns."""
    parser = _sections()
    return parser["DEFAULT"], parser["SIM_FLIGHT"]


def get_batch_config():
    """The ``[BATCH]`` section: ``seed``, ``gain_scale``, ``mass_scale``, ``inertia_scale`` (absent from the reference)."""
    return _sections()["BATCH"]


def parse_array(section, key: str) -> np.ndarray:
    """A list-valued entry such as ``[1, 2, 3]`` as a NumPy array."""
    return np.asarray(json.loads(section.get(key) if hasattr(section, "get") else section[key]))
