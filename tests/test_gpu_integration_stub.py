"""The ctypes stub printed in INTEGRATION.md (what a maintainer of the reference would add) must work as
written: it is extracted from the document, executed, and checked against a golden MCMC trajectory."""
import os
import random
import re

import numpy as np
import pytest

from conftest import ROOT, golden

pytestmark = pytest.mark.gpu


def test_integration_md_stub_reproduces_reference_mcmc():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# nlmc_binding\.py.*?)```", text, flags=re.S).group(1)
    block = block.replace('"nonlocal-monte-carlo_b200/', '"' + os.path.join(ROOT, "nonlocal-monte-carlo_b200") + "/")
    ns = {}
    exec(compile(block, "INTEGRATION.md:nlmc_binding.py", "exec"), ns)
    g = golden("mcmc_element")
    for tag in ("pm_fixed", "gauss_fixed"):
        J, h = g[f"{tag}_J"], g[f"{tag}_h"]
        np.random.seed(int(g[f"{tag}_seed"]))
        random.seed(int(g[f"{tag}_seed"]))
        m0 = np.sign(2 * np.random.rand(len(h)) - 1)
        M = ns["mcmc_gpu"](J, h, m0, np.full(int(g[f"{tag}_sweeps"]), float(g[f"{tag}_beta"])))
        assert M.dtype == np.float64 and np.array_equal(M, g[f"{tag}_M"].astype(float))
