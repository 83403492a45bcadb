"""Shared test cases (the same list the golden generator uses)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import golden_cases  # noqa: E402,F401

ALLOWED = [(4, 4), (4, 5), (4, 8), (4, 32), (8, 8), (8, 9), (8, 10), (8, 32), (8, 64),
           (16, 16), (16, 32), (16, 64), (32, 32), (32, 64), (64, 64)]

# SURVEY.md section 8c fingerprints of config 1 (plan_unordered, 16x16, cos/sin fill rounded to fp32,
# product accumulated in fp64), measured with the reference's formula during the survey
CFG1_FINGERPRINT = dict(sum_re=1.353588395e3, sum_im=-1.090537714e3, sum_abs2=1.453273403e8, max_abs=31.086123,
                        y000=(1.819630876, 21.496855423), y035=(-21.532847687, -1.327526424),
                        ylast=(-1.013397604, 13.188915067))
