"""Generates tests/golden/simple_float.json with the REFERENCE's own perceval/utils/format.py:simple_float (loaded as a
stand-alone file: it imports only sympy and numpy, so it runs in this container although `import perceval` does not).
Run here (the GPU box has no /root/reference):  python tests/golden/make_format_golden.py"""
import importlib.util
import json
import os

import numpy as np

spec = importlib.util.spec_from_file_location("ref_format", "/root/reference/perceval/utils/format.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

rng = np.random.default_rng(0)
vals = [0.0, 1.0, 0.5, 1 / 3, 2 / 3, 1 / 9, 7 / 9, 1e-5, 0.123456789, 0.99999999, 2.5e-7, 3.2e-12, 0.0011, 0.00099, 0.001, 0.01,
        0.1, 0.9999995, 0.9999994, 1e-16, 1.5e-16, 0.38639895265345636, 0.00699, 0.0024, 1e-3 - 1e-9, 0.25, 0.0625, 4.4e-4,
        9.9999996e-5, 9.9999994e-4]
vals += [float(x) for x in rng.random(200)]
vals += [float(10 ** (-12 * x)) for x in rng.random(200)]
vals += [float(x * 10 ** (-k)) for k in range(0, 15) for x in (1.0, 9.9999995, 9.9999994, 1.0000005, 5.5555555)]
out = [[v, mod.simple_float(v, nsimplify=False)[1]] for v in vals]
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "simple_float.json")
json.dump(out, open(path, "w"))
print(len(out), "values ->", path)
