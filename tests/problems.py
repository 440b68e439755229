"""The problem recipes live in the package (mpp_b200/problems.py) so that bench.py, tools/ and smoke() do not depend on the
test tree; the tests keep importing them under this name."""
from mpp_b200.problems import *  # noqa: F401,F403
