"""Drop-in for reference ``web/ml_recommend_web.py``: ``get_ml_recommend(cube_name, amount, root,
non_json) -> {"additions": {card: score}, "cuts": {card: score}}`` (reference :10-67).

Resident serving.  The reference reloads a 390 MB SavedModel on every request (:37); here the model
stays on the GPU, and concurrent requests (Flask ``threaded=True`` / gunicorn threads, reference ``web/__init__.py:41``)
are MICRO-BATCHED: every request thread hands its cube to a ``RequestBatcher`` and blocks on a future; one worker thread
owns the GPU stream, drains whatever is queued (up to ``max_batch`` cubes, waiting at most ``max_wait_ms`` for company
once a first request is in), runs ONE ``MLRecommender.recommend_device`` call over the batch -- the same batched
encoder / decoder / masked-select kernels as BASELINE configs[3] -- and hands every request its slice.  A request's
answer does not depend on what it was batched with (each cube is ranked on its own row of scores).

Set ``CUBECOBRA_MODEL_DIR`` / ``CUBECOBRA_ID_MAP`` to point at the checkpoint and id map; ``CUBECOBRA_MAX_BATCH`` and
``CUBECOBRA_MAX_WAIT_MS`` tune the batcher.
"""
import os
import queue
import threading
import time
from concurrent.futures import Future

import numpy as np

ROOT = "https://cubecobra.com"
_state = {"rec": None, "maps": None, "batcher": None}
_lock = threading.Lock()


class RequestBatcher:
    """Micro-batches concurrent single-cube requests into batched GPU calls.

        batcher = RequestBatcher(MLRecommender(model))
        fut = batcher.submit(cube_indices, amount)        # any thread
        ids, scores, results_at = fut.result()            # additions (ranked), their scores, {in-cube idx: score}

    ``stats`` counts requests, GPU batches and the largest batch, so a test (or an operator) can see the batching."""

    def __init__(self, recommender, max_batch: int = 256, max_wait_ms: float = 2.0):
        self.rec = recommender
        self.max_batch = int(max_batch)
        self.max_wait = float(max_wait_ms) / 1e3
        self.q: "queue.Queue" = queue.Queue()
        self.stats = {"requests": 0, "batches": 0, "max_batch": 0}
        self._stop = False
        self._dead = None                   # the exception that killed the worker, if any
        dev = recommender.model.device      # the worker thread must run on the model's GPU (a new thread starts on device 0)
        self._cuda_index = None
        if dev.type == "cuda":
            import torch
            self._cuda_index = dev.index if dev.index is not None else torch.cuda.current_device()
        self.worker = threading.Thread(target=self._run, name="cubecobra-batcher", daemon=True)
        self.worker.start()

    def submit(self, cube_indices, amount: int) -> Future:
        fut: Future = Future()
        if self._dead is not None or not self.worker.is_alive():
            fut.set_exception(RuntimeError(f"the request batcher's worker thread is not running ({self._dead!r})"))
            return fut
        self.q.put((list(cube_indices), int(amount), fut))
        return fut

    def close(self):
        self._stop = True
        self.q.put(None)
        self.worker.join(timeout=10)

    # -- worker ---------------------------------------------------------------------------------------
    def _collect(self):
        first = self.q.get()
        if first is None:
            return None
        batch = [first]
        deadline = time.monotonic() + self.max_wait
        while len(batch) < self.max_batch:
            left = deadline - time.monotonic()
            try:
                item = self.q.get(timeout=left) if left > 0 else self.q.get_nowait()
            except queue.Empty:
                break
            if item is None:
                self._stop = True
                break
            batch.append(item)
        return batch

    def _run(self):
        batch = []
        try:
            if self._cuda_index is not None:
                import torch
                torch.cuda.set_device(self._cuda_index)
            while not self._stop:
                batch = self._collect()
                if not batch:
                    break
                try:
                    self._serve(batch)
                except Exception as e:  # every waiting request sees the failure (the reference re-raises per request)
                    for _, _, fut in batch:
                        if not fut.done():
                            fut.set_exception(e)
                batch = []
        except BaseException as e:      # the worker itself died: nobody may be left waiting on a future
            self._dead = e
            pending = list(batch or [])
            while True:
                try:
                    item = self.q.get_nowait()
                except queue.Empty:
                    break
                if item is not None:
                    pending.append(item)
            for _, _, fut in pending:
                if not fut.done():
                    fut.set_exception(RuntimeError(f"request batcher worker died: {e!r}"))

    def _serve(self, batch):
        from ..sparse import CubeCSR
        num_cards = self.rec.model.N
        csr = CubeCSR.from_lists([b[0] for b in batch], num_cards)      # duplicates collapse, like cube[idx] = 1
        n_max = max(1, min(max(b[1] for b in batch), num_cards))
        ids, vals, cnts, cuts = self.rec.recommend(csr, n_max, want_cuts=True)
        self.stats["requests"] += len(batch)
        self.stats["batches"] += 1
        self.stats["max_batch"] = max(self.stats["max_batch"], len(batch))
        for r, (idxs, amount, fut) in enumerate(batch):
            take = min(max(amount, 0), int(cnts[r]))
            # (the reference's loop emits one card before testing `recommended >= amount`, so amount <= 0 yields one)
            take = max(take, min(1, int(cnts[r])))
            lo, hi = int(csr.indptr[r]), int(csr.indptr[r + 1])
            at = dict(zip(csr.indices[lo:hi].tolist(), cuts[lo:hi].tolist()))
            fut.set_result((ids[r, :take].copy(), vals[r, :take].copy(), at))


def _resident():
    from ..ml.inference import MLRecommender
    from ..ml.model import load_model
    from ..scripts.common import load_int_to_card
    with _lock:
        if _state["rec"] is None:
            _state["maps"] = load_int_to_card(os.environ.get("CUBECOBRA_ID_MAP", "./ml_files/recommender_id_map.json"))
            _state["rec"] = MLRecommender(load_model(os.environ.get("CUBECOBRA_MODEL_DIR", "./ml_files/recommender")))
        if _state["batcher"] is None:
            _state["batcher"] = RequestBatcher(_state["rec"], int(os.environ.get("CUBECOBRA_MAX_BATCH", "256")),
                                               float(os.environ.get("CUBECOBRA_MAX_WAIT_MS", "2")))
    return _state["batcher"], _state["maps"]


def install(recommender, int_to_card, max_batch=256, max_wait_ms=2.0):
    """Serve from an already-built recommender (tests, embedding applications) instead of loading from disk."""
    with _lock:
        if _state["batcher"] is not None:
            _state["batcher"].close()
        _state["rec"] = recommender
        _state["maps"] = (dict(int_to_card), {v: k for k, v in int_to_card.items()})
        _state["batcher"] = RequestBatcher(recommender, max_batch, max_wait_ms)
    return _state["batcher"]


def format_output(idxs, ids, scores, results_at, int_to_card):
    """The reference's response dict (web/ml_recommend_web.py:48-67): additions in rank order, cuts in cubelist order
    (a dict: a repeated card keeps its first position)."""
    output = {"additions": dict(), "cuts": dict()}
    for rec, score in zip(ids, scores):
        output["additions"][int_to_card[int(rec)]] = float(np.float32(score))
    for idx in idxs:
        output["cuts"][int_to_card[idx]] = float(np.float32(results_at[idx]))
    return output


def get_ml_recommend(cube_name, amount, root=ROOT, non_json=False, card_names=None):
    from ..scripts.common import cube_indices, fetch_cube_list
    if card_names is None:
        card_names = fetch_cube_list(cube_name, root)
    batcher, (int_to_card, card_to_int) = _resident()
    idxs = cube_indices(card_names, card_to_int)
    # blocks this request's thread only; the timeout turns a wedged GPU into an error instead of a hung server thread
    ids, scores, results_at = batcher.submit(idxs, amount).result(timeout=float(os.environ.get("CUBECOBRA_REQUEST_TIMEOUT_S", "600")))
    if non_json:
        for rec in ids:
            print(int_to_card[int(rec)])
        return None
    return format_output(idxs, ids, scores, results_at, int_to_card)
