"""chronomind_b200 — B200-native batched search engine behind ChronoMind's read path.

The product is `libchronomind_b200.so` (C ABI in `include/chronomind_b200.h`, sources in `csrc/`).
This package is the thin host-side mirror of the reference's interfaces over that ABI:
`Index` (PyO3 `Index` precedent, `bindings/python/src/lib.rs:35-123`) and `ChronoMindB200`
(`ChronoMind::search`, `src/store.rs:321-346`). The native library is loaded lazily by
`chronomind_b200.native`; nothing here falls back to a CPU path.
"""
__version__ = "0.1.0"

from . import datasets  # noqa: F401


def __getattr__(name):
    if name == "engine":
        import importlib
        return importlib.import_module(__name__ + ".engine")
    if name in ("Index", "ChronoMindB200", "native", "CmError", "Config", "default_config", "merge_topk_cuda", "rerank_cuda",
                "last_counters", "launch_count", "NO_ID"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
