"""CPU: the collector is held during a graph capture and restored afterwards (million_b200.pq_utils._quiet_gc)."""
import gc


def test_quiet_gc_holds_and_restores_the_collector():
    from million_b200.pq_utils import _quiet_gc
    assert gc.isenabled()
    with _quiet_gc():
        assert not gc.isenabled()
    assert gc.isenabled()
    gc.disable()
    try:
        with _quiet_gc():
            assert not gc.isenabled()
        assert not gc.isenabled()                # left as found
    finally:
        gc.enable()
    try:
        with _quiet_gc():
            raise RuntimeError("x")
    except RuntimeError:
        pass
    assert gc.isenabled()
