"""Minimal stand-in for `hydra` (not installed here, no network): just enough for the reference's entry scripts
(ACT/main_dist.py:34, STH/evaluate.py:28) to be IMPORTED so that their validate() functions can be called directly.
Test infrastructure only (shim 3)."""


def main(config_path=None, config_name=None, **_kw):
    def deco(fn):
        def run(*a, **k):
            raise RuntimeError("hydra stub: call validate()/main_worker() directly with an args namespace")
        run.__wrapped__ = fn
        return run
    return deco
