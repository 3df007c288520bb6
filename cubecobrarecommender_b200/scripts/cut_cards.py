"""Drop-in for reference ``src/scripts/cut_cards.py``.

    python -m cubecobrarecommender_b200.scripts.cut_cards cube_id [N=100]

``simple_cuts(cube, adj_mtx, int_to_card=None)`` (reference cut_cards.py:7-18): in-cube cards ranked
by ascending sum of M[i, j] over the other in-cube cards i != j.  Like the reference (:8) it zeroes the
diagonal of a NumPy ``adj_mtx`` in place; a resident ``GraphRecommender`` is left untouched (the
kernel reads the diagonal as 0 instead).
"""
import sys

import numpy as np

from .recommend import _recommender


def simple_cuts(cube, adj_mtx, int_to_card=None, amount=None):
    from ..sparse import CubeCSR
    if isinstance(adj_mtx, np.ndarray):
        np.fill_diagonal(adj_mtx, 0)                      # cut_cards.py:8 mutates its argument
    cube = np.asarray(cube)
    csr = CubeCSR.from_dense(cube)
    rec = _recommender(adj_mtx)
    n_in = int((cube == 1).sum())
    n = n_in if amount is None else min(int(amount), n_in)
    if n == 0:
        return []
    ids, _, cnt = rec.cuts(csr, n)
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
    recs = simple_cuts(cube, adj_mtx, int_to_card, amount=amount)
    for i in range(amount):
        print(str(i + 1) + ":", recs[i])


if __name__ == "__main__":
    main()
