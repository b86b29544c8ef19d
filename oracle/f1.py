"""CPU restatement of row f1 -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

tagDups (src/pacybara_translate.R:58-68): a barcode is tagged "collision" when `table(bcs)` counts it more than once.
Soft filter (src/pacybara_softfilter.R:46-63): bcIdx = sizes per upBarcode; a row is kept when
size / sum(bcIdx[[upBarcode]]) > minRatio, and afterwards rows with size <= 1 are dropped.
Parity unpinned: R is not available in the build container, so no output of those scripts could be recorded; the
restatement follows the R source line by line and the GPU path is checked against it."""
from collections import Counter
from typing import List, Sequence, Tuple


def tag_table(table: Sequence[Sequence], min_ratio: float = 0.666) -> List[Tuple]:
    bc_count = Counter(r[0] for r in table)
    up_count = Counter(r[1] for r in table)
    up_total = Counter()
    for r in table:
        up_total[r[1]] += int(r[3])
    out = []
    for r in table:
        size = int(r[3])
        usable = size / up_total[r[1]] > min_ratio and size > 1
        out.append(tuple(r) + ("collision" if bc_count[r[0]] > 1 else "", "collision" if up_count[r[1]] > 1 else "", usable))
    return out
