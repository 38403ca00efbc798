"""Host-side logic that needs neither a GPU nor the CUDA library."""


def test_hostmem_numa_helpers_never_raise():
    """bench.py calls bind_to_gpu_numa on every rank: it must degrade to None, not raise, on hosts
    without a GPU, without sysfs NUMA information, or when switched off."""
    import os
    from mri_raytracer_b200 import hostmem
    assert hostmem._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostmem._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    assert hostmem.gpu_numa_node(0) is None or isinstance(hostmem.gpu_numa_node(0), int)
    r = hostmem.bind_to_gpu_numa(0)
    assert r is None or isinstance(r, int)
    os.environ["MRT_NUMA_BIND"] = "0"
    try:
        assert hostmem.bind_to_gpu_numa(0) is None
    finally:
        del os.environ["MRT_NUMA_BIND"]
        os.sched_setaffinity(0, before)
