"""Synthetic barcoded-clone libraries in the formats the clustering hot path consumes.

The shapes follow SURVEY.md section 8(d) / BASELINE.json ``configs``.  A library is
produced as flat numpy arrays (what the C-ABI takes) and can also be written out
as the text files the reference pipeline exchanges:

* ``bcExtract_*.fastq.gz``  -- header ``@<read> pos=<int|NA> len=<int>``
  (reference producer: src/pacybara_extractBCORF.R:150-153,330-335)
* ``genoExtract.csv.gz``    -- ``"<read>","<geno>"``, no header, all quoted
  (reference producer: src/pacybara_extractBCORF.R:338-359)

Everything is driven by one ``numpy.random.Generator`` seed, so a (config, seed)
pair names the workload exactly.
"""
from __future__ import annotations

import gzip
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_INS_SEQS = ["A", "C", "G", "T", "AC", "GT", "ACG", "TTA"]


def mut_string(code: int) -> str:
    """Mutation code -> the HGVS-like key used in genoExtract (extractBCORF.R:346-350)."""
    pos, kind = divmod(int(code), 64)
    pos -= 16  # relpos may be <= 0 (survey 8(b)): codes store pos+16
    if kind < 16:
        return f"{pos}{'ACGT'[kind >> 2]}>{'ACGT'[kind & 3]}"
    if kind < 16 + len(_INS_SEQS):
        return f"{pos}ins{_INS_SEQS[kind - 16]}"
    return f"{pos}del"


def _rand_mut_codes(rng: np.random.Generator, n: int, orf_len: int) -> np.ndarray:
    pos = rng.integers(-3, orf_len + 4, size=n) + 16
    u = rng.random(n)
    ref = rng.integers(0, 4, size=n)
    alt = (ref + rng.integers(1, 4, size=n)) % 4
    kind = np.where(u < 0.9, ref * 4 + alt,
                    np.where(u < 0.95, 16 + rng.integers(0, len(_INS_SEQS), size=n), 63))
    return pos.astype(np.int64) * 64 + kind


def _apply_edits(codes: np.ndarray, lens: np.ndarray, n_edits: np.ndarray,
                 rng: np.random.Generator, p_sub: float, p_ins: float) -> None:
    """In-place random substitutions / insertions / deletions, one per round."""
    n, lmax = codes.shape
    remaining = n_edits.copy()
    cols = np.arange(lmax)[None, :]
    while True:
        idx = np.nonzero(remaining > 0)[0]
        if idx.size == 0:
            break
        remaining[idx] -= 1
        u = rng.random(idx.size)
        kind = np.where(u < p_sub, 0, np.where(u < p_sub + p_ins, 1, 2))
        ln = lens[idx]
        kind = np.where((kind == 1) & (ln >= lmax), 0, kind)   # no room to insert
        kind = np.where((kind == 2) & (ln <= 1), 0, kind)      # nothing left to delete
        # substitutions
        s = idx[kind == 0]
        if s.size:
            p = (rng.random(s.size) * lens[s]).astype(np.int64)
            codes[s, p] = (codes[s, p] + rng.integers(1, 4, size=s.size)) % 4
        # insertions
        s = idx[kind == 1]
        if s.size:
            p = (rng.random(s.size) * (lens[s] + 1)).astype(np.int64)
            src = np.clip(cols - (cols > p[:, None]), 0, lmax - 1)
            rows = np.take_along_axis(codes[s], src, axis=1)
            rows[np.arange(s.size), p] = rng.integers(0, 4, size=s.size)
            codes[s] = rows
            lens[s] += 1
        # deletions
        s = idx[kind == 2]
        if s.size:
            p = (rng.random(s.size) * lens[s]).astype(np.int64)
            src = np.clip(cols + (cols >= p[:, None]), 0, lmax - 1)
            codes[s] = np.take_along_axis(codes[s], src, axis=1)
            lens[s] -= 1


@dataclass
class Library:
    """One synthetic library as arrays (R reads)."""
    name: str
    seed: int
    seq: np.ndarray          # uint8 [R, Lmax] ASCII, zero padded
    qual: np.ndarray         # uint8 [R, Lmax] ASCII Phred+33, zero padded
    length: np.ndarray       # int32 [R]
    geno_ptr: np.ndarray     # int64 [R+1]
    geno_mut: np.ndarray     # int32 [nnz] ids into mut_codes
    geno_q: np.ndarray       # int32 [nnz]
    mut_codes: np.ndarray    # int64 [M] (see mut_string)
    clone_of_read: np.ndarray
    uptag_len: Optional[np.ndarray] = None   # int32 [R]: combo libraries only (uptag = prefix)
    params: Dict = field(default_factory=dict)
    id_style: str = "padded"

    @property
    def n_reads(self) -> int:
        return int(self.length.shape[0])

    def read_id(self, i: int) -> str:
        if self.id_style == "padded":
            return f"m1/{i:09d}/ccs"
        return f"m1/{i}/ccs"          # string order != index order

    def read_ids(self) -> List[str]:
        return [self.read_id(i) for i in range(self.n_reads)]

    def mut_strings(self) -> List[str]:
        return [mut_string(c) for c in self.mut_codes]

    def seq_str(self, i: int) -> str:
        return self.seq[i, : self.length[i]].tobytes().decode()

    def uptag_seq(self) -> np.ndarray:
        """uint8 [R, Lmax] of the uptag part (bcExtract_1) for combo libraries."""
        if self.uptag_len is None:
            return self.seq
        out = self.seq.copy()
        out[np.arange(out.shape[1])[None, :] >= self.uptag_len[:, None]] = 0
        return out

    # ---- text files -------------------------------------------------------
    def geno_str(self, i: int, muts: List[str]) -> str:
        a, b = self.geno_ptr[i], self.geno_ptr[i + 1]
        if a == b:
            return "="
        return ";".join(f"{muts[m]}:{q}" for m, q in zip(self.geno_mut[a:b], self.geno_q[a:b]))

    def write_files(self, outdir: str, compresslevel: int = 1) -> Dict[str, str]:
        import os
        os.makedirs(outdir, exist_ok=True)
        ids = self.read_ids()
        muts = self.mut_strings()
        paths = {}

        def fastq(path, seq, qual, lens, pos):
            with gzip.open(path, "wt", compresslevel=compresslevel) as fh:
                for i in range(self.n_reads):
                    n = int(lens[i])
                    fh.write(f"@{ids[i]} pos={pos} len={n}\n{seq[i, :n].tobytes().decode()}\n+\n"
                             f"{qual[i, :n].tobytes().decode()}\n")

        if self.uptag_len is None:
            paths["bc"] = os.path.join(outdir, "bcExtract_1.fastq.gz")
            fastq(paths["bc"], self.seq, self.qual, self.length, 17)
            paths["uptag"] = paths["bc"]
        else:
            paths["bc"] = os.path.join(outdir, "bcExtract_combo.fastq.gz")
            fastq(paths["bc"], self.seq, self.qual, self.length, "NA")
            paths["uptag"] = os.path.join(outdir, "bcExtract_1.fastq.gz")
            fastq(paths["uptag"], self.seq, self.qual, self.uptag_len, 17)
        paths["geno"] = os.path.join(outdir, "genoExtract.csv.gz")
        with gzip.open(paths["geno"], "wt", compresslevel=compresslevel) as fh:
            for i in range(self.n_reads):
                fh.write(f"\"{ids[i]}\",\"{self.geno_str(i, muts)}\"\n")
        return paths


def make_library(n_clones: int, n_reads: int, bc_len: int = 25, *, seed: int = 1,
                 combo: bool = False, p_err: float = 0.004, p_sub: float = 0.5, p_ins: float = 0.25,
                 p_near: float = 0.02, p_collide: float = 0.0, orf_len: int = 2500,
                 q_clone=(60, 93), q_private=(5, 30), p_private: float = 0.15,
                 p_private_hq: float = 0.01, p_dropout: float = 0.03,
                 base_q=(30, 93), skew: float = 0.0, id_style: str = "padded",
                 name: str = "custom", pattern: Optional[str] = None) -> Library:
    """Generate a library.

    * barcodes are uniform over ACGT (``bc_len`` nt, or 2 x ``bc_len`` when ``combo``), or follow a degenerate
      ``pattern`` repeated to the barcode length: S = C/G, W = A/T, N = any (the reference's template designs its
      barcode as SWSW..., templates/pacybara_parameters.txt:13 -- a DENSE code space: 2^25 codes for 25 nt);
    * ``p_near`` of the clones get a barcode 1-2 substitutions away from another clone
      (half of them with the same genotype) -- this plants the buckets adjacent to two
      clusters that make the reference's row-order dependence observable (SURVEY KAT-3);
    * ``p_collide`` of the clones copy another clone's barcode exactly but carry a disjoint
      genotype (cfg 4: they must stay separate clusters);
    * per-base barcode errors at rate ``p_err`` (sub/ins/del = p_sub/p_ins/rest);
    * genotype: 0-3 mutations per clone, each seen in a read with prob 1-``p_dropout``;
      ``p_private`` of reads carry one private low-quality call, ``p_private_hq`` one
      private high-quality call.
    """
    rng = np.random.default_rng(seed)
    L = bc_len * (2 if combo else 1)
    lmax = min(64, L + 6)
    clone_bc = rng.integers(0, 4, size=(n_clones, L), dtype=np.int64).astype(np.uint8)
    if pattern:
        pat = (pattern * (L // len(pattern) + 1))[:L]
        bit = rng.integers(0, 2, size=(n_clones, L)).astype(np.uint8)
        lo = np.array([{"S": 1, "W": 0}.get(c, 255) for c in pat], dtype=np.uint8)      # codes: A0 C1 G2 T3 -> S = 1 + bit, W = 3 * bit
        is_s, is_w = lo == 1, lo == 0
        clone_bc = np.where(is_s[None, :], 1 + bit, np.where(is_w[None, :], 3 * bit, clone_bc)).astype(np.uint8)

    # clone genotypes (CSR over clones)
    n_mut = rng.integers(0, 4, size=n_clones)
    cg_ptr = np.zeros(n_clones + 1, dtype=np.int64)
    np.cumsum(n_mut, out=cg_ptr[1:])
    cg_codes = _rand_mut_codes(rng, int(cg_ptr[-1]), orf_len)

    def clone_muts(c):
        return cg_codes[cg_ptr[c]:cg_ptr[c + 1]]

    # near neighbours / collisions are derived from a lower-numbered clone
    derived_geno: Dict[int, np.ndarray] = {}
    if n_clones > 1:
        n_near = int(round(p_near * n_clones))
        n_coll = int(round(p_collide * n_clones))
        picks = rng.choice(np.arange(1, n_clones), size=min(n_clones - 1, n_near + n_coll), replace=False)
        for t, c in enumerate(picks):
            src = int(rng.integers(0, c))
            clone_bc[c] = clone_bc[src]
            if t < n_near:
                for _ in range(int(rng.integers(1, 3))):
                    p = int(rng.integers(0, L))
                    clone_bc[c, p] = (clone_bc[c, p] + rng.integers(1, 4)) % 4
                if rng.random() < 0.5:
                    derived_geno[int(c)] = clone_muts(src).copy()
            else:
                # collision: make sure the genotype is non-empty and disjoint
                g = _rand_mut_codes(rng, int(rng.integers(1, 4)), orf_len)
                derived_geno[int(c)] = np.setdiff1d(g, clone_muts(src))
    if derived_geno:
        lists = [derived_geno.get(c, clone_muts(c)) for c in range(n_clones)]
        n_mut = np.array([len(x) for x in lists], dtype=np.int64)
        cg_ptr = np.zeros(n_clones + 1, dtype=np.int64)
        np.cumsum(n_mut, out=cg_ptr[1:])
        cg_codes = np.concatenate(lists) if len(lists) else cg_codes

    # reads -> clones
    if skew > 0:
        w = rng.lognormal(0.0, skew, size=n_clones)
        clone = rng.choice(n_clones, size=n_reads, p=w / w.sum())
    else:
        clone = rng.integers(0, n_clones, size=n_reads)

    codes = np.zeros((n_reads, lmax), dtype=np.uint8)
    codes[:, :L] = clone_bc[clone]
    lens = np.full(n_reads, L, dtype=np.int32)
    uptag_len = None
    if combo:
        # errors hit the two tags independently; the uptag is the prefix of the combo string
        up = np.zeros((n_reads, lmax), dtype=np.uint8)
        dn = np.zeros((n_reads, lmax), dtype=np.uint8)
        up[:, :bc_len] = codes[:, :bc_len]
        dn[:, :bc_len] = codes[:, bc_len:L]
        ul = np.full(n_reads, bc_len, dtype=np.int32)
        dl = np.full(n_reads, bc_len, dtype=np.int32)
        _apply_edits(up[:, :bc_len + 3], ul, rng.binomial(bc_len, p_err, size=n_reads), rng, p_sub, p_ins)
        _apply_edits(dn[:, :bc_len + 3], dl, rng.binomial(bc_len, p_err, size=n_reads), rng, p_sub, p_ins)
        cols = np.arange(lmax)[None, :]
        dn_idx = np.clip(cols - ul[:, None], 0, lmax - 1)
        codes = np.where(cols < ul[:, None], up, np.take_along_axis(dn, dn_idx, axis=1))
        lens = (ul + dl).astype(np.int32)
        uptag_len = ul
    else:
        _apply_edits(codes, lens, rng.binomial(L, p_err, size=n_reads), rng, p_sub, p_ins)
    cols = np.arange(lmax)[None, :]
    valid = cols < lens[:, None]
    seq = np.where(valid, BASES[codes & 3], 0).astype(np.uint8)
    qual = np.where(valid, rng.integers(base_q[0], base_q[1] + 1, size=(n_reads, lmax)) + 33, 0).astype(np.uint8)

    # read genotypes
    n_c = (cg_ptr[1:] - cg_ptr[:-1])[clone]                       # clone calls per read
    r_ptr = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(n_c, out=r_ptr[1:])
    tot = int(r_ptr[-1])
    owner = np.repeat(np.arange(n_reads), n_c)
    within = np.arange(tot) - r_ptr[owner]
    call_code = cg_codes[cg_ptr[clone[owner]] + within]
    keep = rng.random(tot) >= p_dropout
    call_q = rng.integers(q_clone[0], q_clone[1] + 1, size=tot)
    u = rng.random(n_reads)
    has_priv = u < (p_private + p_private_hq)
    priv_hq = u < p_private_hq
    priv_code = _rand_mut_codes(rng, n_reads, orf_len)
    priv_q = np.where(priv_hq, rng.integers(40, 70, size=n_reads),
                      rng.integers(q_private[0], q_private[1] + 1, size=n_reads))
    all_owner = np.concatenate([owner[keep], np.nonzero(has_priv)[0]])
    all_code = np.concatenate([call_code[keep], priv_code[has_priv]])
    all_q = np.concatenate([call_q[keep], priv_q[has_priv]])
    order = np.argsort(all_owner, kind="stable")
    all_owner, all_code, all_q = all_owner[order], all_code[order], all_q[order]
    # a private call that repeats one of the read's own clone calls would be a duplicate key
    key = all_owner.astype(np.int64) * (1 << 40) + all_code
    _, first = np.unique(key, return_index=True)
    sel = np.sort(first)
    all_owner, all_code, all_q = all_owner[sel], all_code[sel], all_q[sel]
    mut_codes, geno_mut = np.unique(all_code, return_inverse=True)
    geno_ptr = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(np.bincount(all_owner, minlength=n_reads), out=geno_ptr[1:])

    return Library(name=name, seed=seed, seq=seq, qual=qual, length=lens, geno_ptr=geno_ptr,
                   geno_mut=geno_mut.astype(np.int32), geno_q=all_q.astype(np.int32),
                   mut_codes=mut_codes.astype(np.int64), clone_of_read=clone.astype(np.int32),
                   uptag_len=uptag_len, id_style=id_style,
                   params=dict(n_clones=n_clones, n_reads=n_reads, bc_len=bc_len, combo=combo,
                               p_err=p_err, p_near=p_near, p_collide=p_collide))


# The named workloads of BASELINE.json `configs` (clustering flags alongside).
CONFIGS = {
    "cfg1": dict(gen=dict(n_clones=10_000, n_reads=50_000, bc_len=25), max_diff=2, min_qual=32),
    "cfg2": dict(gen=dict(n_clones=100_000, n_reads=1_000_000, bc_len=25), max_diff=2, min_qual=32),
    "cfg3": dict(gen=dict(n_clones=500_000, n_reads=5_000_000, bc_len=25, combo=True), max_diff=3, min_qual=32),
    "cfg4": dict(gen=dict(n_clones=1_000_000, n_reads=10_000_000, bc_len=25, p_err=0.001,
                          p_collide=0.05, q_clone=(20, 50), q_private=(3, 15)), max_diff=2, min_qual=27),
    "cfg5": dict(gen=dict(n_clones=2_000_000, n_reads=20_000_000, bc_len=25), max_diff=2, min_qual=32),
    # not a BASELINE.json config: the reference template's SWSW... barcode design, where 10^6 clones sit in 2^25 codes and
    # the link graph percolates into giant components (hard part 1)
    "dense": dict(gen=dict(n_clones=1_000_000, n_reads=10_000_000, bc_len=25, pattern="SW"), max_diff=2, min_qual=32),
}


def make_config(name: str, seed: int = 1, scale: float = 1.0, **over) -> Library:
    cfg = CONFIGS[name]
    gen = dict(cfg["gen"])
    gen.update(over)
    if scale != 1.0:
        gen["n_clones"] = max(1, int(gen["n_clones"] * scale))
        gen["n_reads"] = max(1, int(gen["n_reads"] * scale))
    lib = make_library(seed=seed, name=name, **gen)
    lib.params.update(max_diff=cfg["max_diff"], min_qual=cfg["min_qual"], min_jaccard=0.2, min_matches=1)
    return lib
