"""Configuration helpers, same functions and keys as the reference ``uav_ac/utils.py:8-28``."""
import ast
import configparser
from pathlib import Path

import numpy as np


def get_config():
    """:return: config object (default, flight) -- ``uav_ac/utils.py:8-19``; the batched additions live in ``[BATCH]``."""
    config = configparser.ConfigParser(inline_comment_prefixes="#")
    config.read(Path(Path(__file__).parent, "config.ini"))
    return config["DEFAULT"], config["SIM_FLIGHT"]


def get_batch_config():
    """The ``[BATCH]`` section (Monte-Carlo seed and perturbation ranges); not present in the reference."""
    config = configparser.ConfigParser(inline_comment_prefixes="#")
    config.read(Path(Path(__file__).parent, "config.ini"))
    return config["BATCH"]


def parse_array(section: configparser.SectionProxy, key: str) -> np.ndarray:
    """Entry holding a Python list literal -> numpy array (``uav_ac/utils.py:22-28``)."""
    return np.array(ast.literal_eval(section.get(key)))
