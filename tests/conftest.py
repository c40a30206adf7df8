import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """A kernel that wedges must end the run, not hang it: with pytest-timeout installed every GPU test gets 15 minutes (the
    longest takes seconds); method "thread" because a test stuck inside a CUDA call never returns to the interpreter."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(900, method="thread"))


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The native libraries are built in-tree once per session (CPU-only cross compile works)."""
    from beom_b200 import build
    from oracle import build as obuild
    build.build_all()
    obuild.build_oracle()
    # tests/test_emulation.py re-runs the GPU test modules on the CPU emulation of the library (tools/emu, test
    # infrastructure): it starts pytest again with BEOM_TEST_EMU naming the emulated library, which is swapped in here
    emu = os.environ.get("BEOM_TEST_EMU")
    if emu:
        import ctypes
        from beom_b200 import _lib
        lib = _lib.bind_gpu(ctypes.CDLL(emu, mode=ctypes.RTLD_LOCAL))
        assert b"cpu-emulation" in lib.beom_gpu_version()
        _lib.host_lib()
        _lib._gpu = lib


# small versions of the named configs: same code paths, sizes the CPU oracle finishes in seconds
SMALL = {
    "stommel1948": dict(dl=250.0e3),
    "lock_exchange": dict(),
    "unstable_jet": dict(dl=60.0e3),
    "sill_exchange3D": dict(lx=6.0e3, ly=100.0e3),
    "conservation": dict(dl=30.0e3),
    "soliton": dict(dl=80.0e3),
    "baines_ridge": dict(scale=0.3),
    "carrier_beach": dict(mesh=20.0),
    "upwelling_seaward_wind": dict(lm=80),
    "mixed_open_bc": dict(lm=60, mm=40),
    "morel_upwelling": dict(dl=4.0e3),
    "outcrop_seamount": dict(),
    "sill_exchange2D": dict(lx=140.0e3),
    "sill_exchange2Dtides": dict(lx=60.0e3),
    "tide_ridge": dict(lm=200),
    "wave_sponge": dict(dl=20.0e3),
    "random_coast": dict(seed=3, lm=150, mm=90),
}


@pytest.fixture(scope="session")
def case_factory(tmp_path_factory):
    from beom_b200 import cases, model

    cache = {}

    def make(name, small=True, variant=0, extra=None, **kw):
        key = (name, small, variant, tuple(sorted((extra or {}).items())), tuple(sorted(kw.items())))
        if key not in cache:
            args = dict(SMALL[name]) if (small and name in SMALL) else {}
            args.update(kw)
            c = cases.CASES[name](**args)
            if extra:
                c.params_text += "".join("%-10s = %s\n" % kv for kv in extra.items())
            d = str(tmp_path_factory.mktemp(name))
            blk = c.write(d)
            cache[key] = (c, d, blk)
        c, d, blk = cache[key]
        hm = model.HostModel.from_block(blk, variant=variant)
        return c, d, hm

    return make
