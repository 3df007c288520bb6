"""WSGI app with the reference's route (reference ``web/__init__.py:16-37``):

    GET /?cube_name=<id>&num_recs=<int, default 30000>&root=<url>   ->  JSON {"additions", "cuts"}

Flask is not a dependency here; ``app`` is a plain WSGI callable (``gunicorn
cubecobrarecommender_b200.web:app``) that returns the same plain-text error strings and the same JSON
(keys sorted, like Flask 1.1's jsonify)."""
import json
import logging
from urllib.parse import parse_qs

from .ml_recommend_web import get_ml_recommend

logger = logging.getLogger("cubecobra_b200.web")


def _text(start_response, body, status="200 OK", ctype="text/html; charset=utf-8"):
    data = body.encode("utf-8")
    start_response(status, [("Content-Type", ctype), ("Content-Length", str(len(data)))])
    return [data]


def app(environ, start_response):
    q = parse_qs(environ.get("QUERY_STRING", ""))
    cube_name = (q.get("cube_name") or [None])[0]
    num_recs = (q.get("num_recs") or [30000])[0]
    root = (q.get("root") or ["https://www.cubecobra.com"])[0]
    if not (cube_name and num_recs):
        error = "Need cube_name and num_recs as parameters!"
        logger.error(error)
        return _text(start_response, error)
    try:
        num_recs = int(num_recs)
    except ValueError:
        error = "num_recs needs to be an integer!"
        logger.error(error)
        return _text(start_response, error)
    try:
        results = get_ml_recommend(cube_name, num_recs, root)
    except Exception as e:  # log and re-raise, like the reference
        logger.error(e)
        raise
    return _text(start_response, json.dumps(results, sort_keys=True), ctype="application/json")
