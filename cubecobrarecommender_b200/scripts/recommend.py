"""Drop-in for reference ``src/scripts/recommend.py``.

    python -m cubecobrarecommender_b200.scripts.recommend cube_id [N=100]

``simple_recs(cube, adj_mtx, int_to_card=None)`` keeps the reference signature (reference
recommend.py:7-18): the full descending ranking of the cards missing from the cube.  Scores are the
float64 column sums in NumPy's pairwise order (bit-identical to the reference); ties are ordered
larger-index-first (``argsort(kind='stable')[::-1]``; the reference's default sort leaves tie order
unspecified).
"""
import sys

import numpy as np


def _recommender(adj_mtx):
    import torch
    from ..graph import GraphRecommender
    if isinstance(adj_mtx, GraphRecommender):
        return adj_mtx
    m = adj_mtx if isinstance(adj_mtx, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(adj_mtx, dtype=np.float64))
    return GraphRecommender(m.to("cuda"))


def simple_recs(cube, adj_mtx, int_to_card=None, amount=None):
    """``adj_mtx``: float64 (C, C) ndarray (uploaded per call, like the reference's np.load per
    invocation) or a resident ``GraphRecommender``."""
    from ..sparse import CubeCSR
    cube = np.asarray(cube)
    csr = CubeCSR.from_dense(cube)
    rec = _recommender(adj_mtx)
    n_missing = int((cube == 0).sum())
    n = n_missing if amount is None else min(int(amount), n_missing)
    if n == 0:
        return []
    ids, _, cnt = rec.recs(csr, n)
    rec_ids = [int(i) for i in ids[0, :int(cnt[0])].cpu().numpy()]
    if int_to_card is None:
        return rec_ids
    return [int_to_card[i] for i in rec_ids]


def main(argv=None):
    from .common import cube_indices, cube_vector, fetch_cube_list, load_int_to_card
    args = sys.argv[1:] if argv is None else argv
    cube_name = args[0]
    amount = int(args[1]) if len(args) > 1 else 100
    print('Getting Cube List . . . \n')
    card_names = fetch_cube_list(cube_name)
    print('Loading Adjacency Matrix . . . \n')
    adj_mtx = np.load('././output/full_adj_mtx.npy')
    print('Loading Card Name Lookup . . . \n')
    int_to_card, card_to_int = load_int_to_card('././output/int_to_card.json')
    print('Creating Cube Vector . . . \n')
    cube = cube_vector(cube_indices(card_names, card_to_int), adj_mtx.shape[1])
    print('Generating Recommendations . . . \n')
    recs = simple_recs(cube, adj_mtx, int_to_card, amount=amount)
    for i in range(amount):
        print(str(i + 1) + ":", recs[i])


if __name__ == "__main__":
    main()
