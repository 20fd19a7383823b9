"""Host logic of the tensor-memory scatter (no GPU): the range plan both kernels derive from the level shapes
(plan_ranges / coarse_first_level in csrc/msda_common.cuh) is compiled for the HOST with nvcc and checked exhaustively --
every owned pixel in exactly one range, the scatter kernel's pixel -> range formula consistent with it, level lists exact."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("plan") / "plan_ranges_check")
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-w", "-o", exe, os.path.join(ROOT, "tests", "host", "plan_ranges_check.cu")])
    return exe


CASES = [
    # max_levels, shapes, expected (first_level, nranges, first level of the first-generation tail)
    (16, [(100, 167), (50, 84), (25, 42), (13, 21)], (0, 31, 2)),          # Swin-T 800x1333: 22 + 6 + 2 + 1 ranges
    (2, [(100, 167), (50, 84), (25, 42), (13, 21)], (2, 3, 2)),
    (3, [(100, 167), (50, 84), (25, 42), (13, 21)], (1, 9, 2)),
    (16, [(256, 450), (128, 225), (64, 113), (32, 57), (16, 29)], (1, 52, 4)),   # Swin-B stride 4: level 0 would need 150 ranges
    (16, [(128, 225), (64, 113), (32, 57), (16, 29), (8, 15)], (0, 52, 3)),
    (16, [(20, 30), (10, 15), (5, 8), (3, 4), (2, 2)], (0, 2, 1)),         # four small levels merge, the fifth gets its own range
    (16, [(48, 40), (9, 7)], (0, 4, 1)),
    (16, [(3, 5)], (0, 1, 0)),
    (16, [(1, 1)], (0, 1, 0)),
    (1, [(60, 70), (30, 35), (15, 18), (8, 9)], (3, 1, 1)),
    (16, [(24, 32), (24, 32), (24, 32)], (0, 3, 1)),                       # 768-pixel levels: exactly one full range each
    (16, [(769, 1)], (0, 2, 0)),
]


@pytest.mark.parametrize("max_levels,shapes,want", CASES)
def test_plan_ranges(checker, max_levels, shapes, want):
    args = [checker, str(max_levels)] + [str(v) for hw in shapes for v in hw]
    out = subprocess.run(args, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    got = json.loads(out.stdout)
    assert got["bad"] == 0
    assert (got["first_level"], got["nranges"], got["tail_first"]) == want
