"""Times the UNMODIFIED reference ``utils.create_adjacency_matrix`` (/root/reference/src/non_ml/utils.py:75-92) on the
configs[0] input (K = 20 000 synthetic cubes x C = 21 000 cards) in the BUILD container (the reference tree does not travel
to the GPU box, and bench.py may not read it there).  Single-threaded by construction.  Writes one JSON line.

    python profiles/reference_create_adjacency_cpu.py [K] > profiles/r02_reference_create_adjacency_cpu.json
"""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference/src/non_ml")
import utils  # noqa: E402  (the reference's own module)

from cubecobrarecommender_b200.workload import make_cubes  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
C = 21000
csr = make_cubes(20000, C, cfg=1).rows(np.arange(K))
cubes = csr.to_dense(np.float64)
t0 = time.time()
adj = utils.create_adjacency_matrix(cubes, verbose=False)
dt = time.time() - t0
from oracle import graph as og  # noqa: E402
same = bool(np.array_equal(adj, og.adjacency_from_counts(og.cooc_counts_blocked(csr.indptr, csr.indices, C))))
print(json.dumps({"what": "unmodified reference utils.create_adjacency_matrix", "K": K, "C": C, "nnz": int(csr.indptr[-1]),
                  "seconds": dt, "threads": 1, "host": f"build container, {os.cpu_count()} vCPU",
                  "numpy": np.__version__, "equals_oracle_bit_for_bit": same}))
