"""hostbind: core-list parsing, and that a rank which binds nothing says why (CPU only)."""
from basecount_b200 import hostbind


def test_cpulist_parsing():
    assert hostbind._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostbind._parse_cpulist("") == set()
    assert hostbind._parse_cpulist("5") == {5}


def test_bind_reports_a_reason_when_it_binds_nothing():
    info = hostbind.bind_to_device_node(0)
    assert set(info) >= {"pci", "node", "cpus", "bound", "reason", "nodes_online"}
    if not info["bound"]:
        assert isinstance(info["reason"], str) and info["reason"]
    else:
        assert info["cpus"] and info["reason"] is None
