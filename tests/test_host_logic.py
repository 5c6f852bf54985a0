"""CPU: host-side logic that mirrors the reference interface — names, heuristics, singletons, page allocation."""
import pytest
import torch

from million_b200 import bindings, pq_utils
from million_b200.paged_pq_utils import PageManager


def test_l2Ns_and_nbits_match_golden(golden):
    assert [pq_utils.l2Ns(int(l)) for l in golden["l2Ns_l"]] == list(golden["l2Ns_v"])
    sizes = [torch.empty(0, dtype=pq_utils.nbits2dtype(int(b))).element_size() for b in golden["nbits"]]
    assert sizes == list(golden["nbits_itemsize"])
    with pytest.raises(ValueError):
        pq_utils.nbits2dtype(65)
    assert pq_utils.scalarTypeToStr(torch.float16) == "f16" and pq_utils.scalarTypeToStr(torch.float32) == "f32"
    with pytest.raises(ValueError):
        pq_utils.scalarTypeToStr(torch.int8)


def test_kernel_name_grammar():
    # the reference's 240-name grid (setup.py:10-15) plus bf16 and the paged variant all resolve
    for sc in ("f16", "bf16"):
        for Ns in (2, 4, 8, 16, 32):
            for d in (64, 128):
                for M in (16, 32, 64):
                    for C in (128, 256):
                        assert callable(getattr(bindings, f"flash_decoding_allocated_buffer_{sc}u8_Ns{Ns}Lt{d}d{d}M{M}C{C}"))
    assert callable(bindings.flash_decoding_paged_v_f16u8_Ns32Lt128d128M64C256)
    for bad in ("flash_decoding_allocated_buffer_f32u8_Ns2Lt64d64M16C128", "flash_decoding", "flash_decoding_allocated_buffer_f16u8_Ns2Lt64d64M48C128"):
        with pytest.raises(AttributeError):
            getattr(bindings, bad)
    import sys
    mod = bindings.install()
    assert sys.modules["bindings"] is mod and __import__("bindings") is mod
    del sys.modules["bindings"]


def test_registry_rejects_what_the_reference_rejects():
    reg = pq_utils.KernelRegistry(M=64, d=128, nbits=9, nh=32, scalar_t=torch.float16)
    with pytest.raises(NotImplementedError):
        reg.get_kernel(4096)
    reg = pq_utils.KernelRegistry(M=64, d=128, nbits=8, nh=32, scalar_t=torch.float32)   # f32 kernels are not compiled (setup.py:9)
    with pytest.raises(AttributeError):
        reg.get_kernel(4096)


def test_singleton_is_per_class():
    class A(metaclass=pq_utils.Singleton):
        def __init__(self, x=0):
            self.x = x

    class B(metaclass=pq_utils.Singleton):
        pass

    pq_utils.Singleton.clear_instance()
    assert not A.has_instance()
    a = A(3)
    assert A(5) is a and a.x == 3 and A.has_instance() and not B.has_instance()
    pq_utils.Singleton.clear_instance()
    assert not A.has_instance()


def test_page_manager_allocation_order_growth_and_limits():
    pm = PageManager(page_size=64, initial_pages=4, max_pages=None, M=64, device="cpu")
    assert pm.page_pool.shape == (8, 64, 64)
    ids = pm.allocate_pages(3)
    assert ids == [0, 1, 2]                       # ascending from a fresh pool
    pm.free_page(1)
    assert pm.allocate_page() == 1                # lowest id first, counted as re-use
    assert pm.page_reuse_count == 1
    more = pm.allocate_pages(10)                  # forces activation + re-allocation
    assert more == list(range(3, 13)) and pm.page_pool.shape[0] >= 13 and pm.total_expansions >= 1
    assert pm.get_stats()["allocated_pages"] == 13
    with pytest.raises(ValueError):
        pm.get_page(10_000)
    capped = PageManager(page_size=64, initial_pages=2, max_pages=3, M=64, device="cpu")
    capped.allocate_pages(3)
    with pytest.raises(RuntimeError):
        capped.allocate_page()


def test_page_manager_bulk_allocation_matches_sequential():
    """allocate_pages(n) must hand out exactly what n allocate_page() calls would (lowest id first, reuse counted, pool grown),
    dynamic_paged_pq_utils.py:65-135; PageManager is plain host logic and runs on a CPU pool."""
    from million_b200.paged_pq_utils import PageManager
    a = PageManager(64, initial_pages=8, M=4, device='cpu')
    b = PageManager(64, initial_pages=8, M=4, device='cpu')
    got = a.allocate_pages(5) + a.allocate_pages(9)            # the second call has to grow the pool
    want = [b.allocate_page() for _ in range(14)]
    assert got == want == list(range(14))
    for pm in (a, b):
        pm.free_page(3); pm.free_page(11); pm.free_page(0)
    assert a.allocate_pages(4) == [b.allocate_page() for _ in range(4)] == [0, 3, 11, 14]
    sa, sb = a.get_stats(), b.get_stats()
    for k in ('allocated_pages', 'page_reuse_count', 'total_allocations'):      # pool growth differs by design (bulk grows by exactly what it needs)
        assert sa[k] == sb[k], k
    assert a.allocated_pages[3]['allocation_count'] == 1 and sa['page_reuse_count'] == 3
    assert a.allocate_pages(0) == []
    capped = PageManager(64, initial_pages=4, max_pages=6, M=4, device='cpu')
    capped.allocate_pages(6)
    with pytest.raises(RuntimeError):
        capped.allocate_pages(1)
