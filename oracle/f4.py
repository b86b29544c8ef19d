"""CPU restatement of row f4 -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

src/pacybara_merge.R:48-116, line by line: index the first library's rows by virtualBarcode (:50-51); for each row of
the second library (:59-86) -- barcode unknown: new, to be appended (:66-69); a row of the first library has the same
hgvsc: remember (row, those rows) (:72-78); else: a conflict, appended as a row of its own (:79-83).  Then the merge loop
(:94-103) in the order the pairs were recorded: of several candidate rows the one with the largest CURRENT size
(which.max: the first of equals), whose size grows by the second row's and whose read list gets the second row's
appended with "|".  Then rbind the new and conflicting rows (:106) and re-tag collisions on virtualBarcode and upBarcode
(:110-116).  Parity unpinned: R is not available in the build container, so no output of the script could be recorded."""
from collections import Counter
from typing import Dict, List, Sequence


def merge_tables(lib1: Sequence[Dict], lib2: Sequence[Dict]) -> List[Dict]:
    """Rows are dicts with virtualBarcode, upBarcode, reads, size, hgvsc (+ whatever else)."""
    idx: Dict[str, List[int]] = {}
    for i, r in enumerate(lib1):
        idx.setdefault(r["virtualBarcode"], []).append(i)
    to_merge, to_add = [], []
    for row, r in enumerate(lib2):
        hits = idx.get(r["virtualBarcode"])
        if hits is None:
            to_add.append(row)
        elif any(lib1[h]["hgvsc"] == r["hgvsc"] for h in hits):
            to_merge.append((row, [h for h in hits if lib1[h]["hgvsc"] == r["hgvsc"]]))
        else:
            to_add.append(row)
    out = [dict(r) for r in lib1]
    for row, hit in to_merge:
        h = hit[0]
        for c in hit[1:]:
            if out[c]["size"] > out[h]["size"]:
                h = c
        out[h]["size"] = out[h]["size"] + lib2[row]["size"]
        out[h]["reads"] = out[h]["reads"] + "|" + lib2[row]["reads"]
    out += [dict(lib2[row]) for row in to_add]
    for col, tag in (("virtualBarcode", "collision"), ("upBarcode", "upTagCollision")):
        cnt = Counter(r[col] for r in out)
        for r in out:
            r[tag] = "collision" if cnt[r[col]] > 1 else ""
    return out
