"""Drop-in for reference ``src/scripts/similarity.py``.

    python -m cubecobrarecommender_b200.scripts.similarity card_name N

``card_name`` with underscores for spaces (reference similarity.py:8).  Prints ``"{rank}: {name} {dist}"`` for the
``N`` nearest cards by the Keras ``CosineSimilarity`` loss (= -cosine) between encoder embeddings, most similar first
(the card itself leads with -1.0), like similarity.py:31-35.  The reference encodes a dense C x C identity and calls the
loss C times in a Python loop; here the embeddings come from the embedding-bag encoder and one kernel computes all C
distances.
"""
import sys


def main(argv=None, model_dir='ml_files/high_req', id_map='ml_files/recommender_id_map.json'):
    from ..ml.inference import MLRecommender
    from ..ml.model import load_model
    from .common import load_int_to_card
    args = sys.argv[1:] if argv is None else argv
    name = args[0].replace('_', ' ')
    n = int(args[1])
    int_to_card, card_to_int = load_int_to_card(id_map)
    rec = MLRecommender(load_model(model_dir))
    ids, dists = rec.similar(card_to_int[name], n)
    out = []
    for i, (card_idx, d) in enumerate(zip(ids, dists)):
        print(str(i + 1) + ":", int_to_card[int(card_idx)], float(d))
        out.append((int_to_card[int(card_idx)], float(d)))
    return out


if __name__ == "__main__":
    main()
