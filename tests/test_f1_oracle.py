"""Row f1 restatement (oracle/f1.py) on hand-checkable tables (R is unavailable: parity unpinned, see the module)."""
from oracle import f1
from tests import util


def test_kat1_collision_survives():
    """KAT-1: r6 carries the same barcode as the 4-read cluster with another genotype: both rows are tagged."""
    case = util.load_golden("kat1")
    table = [(r[1], r[2], r[0], r[3], r[4]) for r in util.golden_table(case)]
    tagged = {r[2]: r[5:] for r in f1.tag_table(table)}
    assert tagged["r1|r2|r3|r4"] == ("collision", "collision", True)        # 4 of 5 reads of that uptag: dominant
    assert tagged["r6"] == ("collision", "collision", False)                # size 1 is never usable
    assert tagged["r5"] == ("", "", False)


def test_dominance_is_strict_and_needs_two_reads():
    table = [("A", "U", "a", 2, "="), ("B", "U", "b", 1, "="), ("C", "V", "c", 3, "="), ("D", "W", "d", 1, "=")]
    out = f1.tag_table(table, 0.666)
    assert [r[7] for r in out] == [True, False, True, False]                # 2/3 = 0.6667 > 0.666; 3/3; size 1 dropped
    assert [r[7] for r in f1.tag_table(table, 2 / 3)] == [False, False, True, False]   # strict >
    assert [r[6] for r in out] == ["collision", "collision", "", ""]
