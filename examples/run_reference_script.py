#!/usr/bin/env python
"""Run a script written for the reference package (`from prosstt import ...`) on prosstt_b200.

    python examples/run_reference_script.py /path/to/generate_simN.py -j job -o outdir -n 2

The package is registered under the name `prosstt` (prosstt_b200.install_as_prosstt) and the
plotting-only imports of the reference example (matplotlib/pylab/anndata/scanpy, absent on
compute nodes) are stubbed, then the script runs unchanged under runpy."""
import os
import runpy
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prosstt_b200  # noqa: E402

prosstt_b200.install_as_prosstt()
for name in ("matplotlib", "matplotlib.pyplot", "pylab", "anndata", "scanpy", "scanpy.api", "scanpy.api.tl"):
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            stub = types.ModuleType(name)
            stub.diffmap = lambda *a, **k: None
            sys.modules[name] = stub
            parent, _, child = name.rpartition(".")
            if parent and parent in sys.modules:
                setattr(sys.modules[parent], child, stub)

if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    runpy.run_path(script, run_name="__main__")
