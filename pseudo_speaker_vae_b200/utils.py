"""The integer tables and parsers on the sampling path (bit-exact rows of SURVEY 8 a11).

Label tables: ps_vae/utils.py:82-136 (Common Voice gender / age buckets, VCTK gender; unknown -> -1).
Target parsing: ps_vae/inference.py:128-132.  Output indexing: ps_vae/inference.py:154-156.
"""
from __future__ import annotations

import json
from typing import Union

_CV_AGE = {"teens": 0, "twenties": 0, "thirties": 1, "fourties": 1, "fifties": 1, "sixties": 2, "seventies": 2, "eighties": 2, "nineties": 2}
_CV_GENDER = {"male": 0, "female": 1, "other": 2}
_VCTK_GENDER = {"M": 0, "F": 1}


def map_cv_age_to_label(age) -> int:
    return _CV_AGE.get(age, -1)


def map_cv_gender_to_label(gender) -> int:
    return _CV_GENDER.get(gender, -1)


def map_vctk_gender_to_label(gender) -> int:
    return _VCTK_GENDER.get(gender, -1)


def parse_classifier_target(text: str) -> Union[int, dict]:
    """``--classifier_target`` of the inference CLI: JSON (a dict of label -> class) first, else an int."""
    try:
        return json.loads(text)
    except json.JSONDecodeError:
        return int(text)


def sample_filename(i: int) -> str:
    """Row ``i`` of a sampled batch is saved as ``sample_{i}.pt``."""
    return f"sample_{i}.pt"


def load_yaml_config(path: str) -> dict:
    """ps_vae/utils.py:69-80."""
    import yaml

    with open(path, "r") as f:
        return yaml.safe_load(f)
