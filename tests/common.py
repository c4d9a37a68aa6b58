"""Shared helpers of the test-suite. The generators themselves live in the package
(diffusionremotesensing_b200/synthetic.py) because bench.py and smoke() use them too."""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from diffusionremotesensing_b200.synthetic import (FAMILIES, build_model, default_init_model, max_rel_err,  # noqa: E402,F401
                                                   np_rand, np_randn, psnr_ref_range, synthetic_state_dict)
