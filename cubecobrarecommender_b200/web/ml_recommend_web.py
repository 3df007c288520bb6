"""Drop-in for reference ``web/ml_recommend_web.py``: ``get_ml_recommend(cube_name, amount, root,
non_json) -> {"additions": {card: score}, "cuts": {card: score}}`` (reference :10-67).

The model stays resident on the GPU between requests (the reference reloads the SavedModel on every
call, :37); set ``CUBECOBRA_MODEL_DIR`` / ``CUBECOBRA_ID_MAP`` to point at the checkpoint and id map.
"""
import os
import threading

ROOT = "https://cubecobra.com"
_state = {"rec": None, "maps": None}
_lock = threading.Lock()


def _resident():
    from ..ml.inference import MLRecommender
    from ..ml.model import load_model
    from ..scripts.common import load_int_to_card
    with _lock:
        if _state["rec"] is None:
            _state["maps"] = load_int_to_card(os.environ.get("CUBECOBRA_ID_MAP", "./ml_files/recommender_id_map.json"))
            _state["rec"] = MLRecommender(load_model(os.environ.get("CUBECOBRA_MODEL_DIR", "./ml_files/recommender")))
    return _state["rec"], _state["maps"]


def get_ml_recommend(cube_name, amount, root=ROOT, non_json=False, card_names=None):
    from ..scripts.common import cube_indices, fetch_cube_list
    if card_names is None:
        card_names = fetch_cube_list(cube_name, root)
    rec, (int_to_card, card_to_int) = _resident()
    idxs = cube_indices(card_names, card_to_int)
    with _lock:                                   # one GPU stream; requests are serialised
        output = rec.recommend_one(idxs, amount, int_to_card)
    if non_json:
        for card in output["additions"]:
            print(card)
        return None
    return output
