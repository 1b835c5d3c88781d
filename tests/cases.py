"""Seeded CSR test shapes shared by the CPU (oracle) and GPU (parity) tests.

The reference has no test suite; these are the adversarial shapes its algorithm
distinguishes: empty rows (dirty tiles + offset table), rows much longer than a tile (fast-track
tiles, multi-tile carries), p = 1 (tail only), nnz % (omega*sigma) == 0, trailing / leading empty
rows, two-packet descriptors (sigma 26, 32), m = 1, and the reference's own auto-sigma choices.
Each case is (name, HostCsr, sigma) with sigma = -1 meaning ANONYMOUSLIB_AUTO_TUNED_SIGMA.
"""
import numpy as np

from benchmark_spmv_using_csr5_b200 import matrices as M


def _counts(rng, m, lo, hi):
    return rng.integers(lo, hi, size=m)


def small_cases():
    rng = np.random.default_rng(7)
    out = []
    out.append(("banded16_s16", M.banded(2000, 16), 16))
    out.append(("banded16_auto", M.banded(4096, 16), -1))
    out.append(("banded16_exact_multiple", M.banded(32 * 16 * 3 // 16, 16), 16))  # nnz = 3 tiles exactly
    out.append(("random_noempty_s8", M.from_row_counts(_counts(rng, 3000, 1, 17), 3000, 1), 8))
    c = _counts(rng, 5000, 0, 9)
    c[1000] = 5000
    out.append(("empty_rows_long_row_s4", M.from_row_counts(c, 5000, 2), 4))
    out.append(("example_c1_auto", M.example_c1(), -1))
    out.append(("two_packet_s32", M.from_row_counts(_counts(rng, 1500, 20, 60), 1500, 3), 32))
    c = _counts(rng, 1500, 0, 60)
    out.append(("two_packet_s26_empty", M.from_row_counts(c, 1500, 4), 26))
    c = _counts(rng, 2000, 1, 12)
    c[-37:] = 0
    out.append(("trailing_empty_s5", M.from_row_counts(c, 2000, 5), 5))
    c = _counts(rng, 2000, 1, 12)
    c[:41] = 0
    out.append(("leading_empty_s7", M.from_row_counts(c, 2000, 6), 7))
    out.append(("p1_tiny", M.from_row_counts([1, 0, 2], 5, 7), 4))
    out.append(("m1_one_row", M.from_row_counts([1000], 50, 8), 4))
    out.append(("m1_short", M.from_row_counts([3], 50, 9), 16))
    c = np.zeros(777, np.int64)
    c[::3] = 9
    out.append(("mostly_empty_s6", M.from_row_counts(c, 777, 10), 6))
    # trailing empty rows AND nnz a multiple of the tile: reference hazard App. B (OOB atomicOr)
    c = np.full(64, 8, np.int64)
    c = np.concatenate([c, np.zeros(9, np.int64)])
    out.append(("trailing_empty_exact_multiple", M.from_row_counts(c, 100, 11), 4))
    # one huge row in the middle spanning many tiles, neighbours short, some empty
    c = _counts(rng, 600, 0, 5)
    c[300] = 20000
    out.append(("hub_row_s12", M.from_row_counts(c, 4000, 12), 12))
    # rows starting exactly on tile boundaries (every row = one tile)
    out.append(("rows_eq_tile_s4", M.from_row_counts(np.full(40, 128), 999, 13), 4))
    out.append(("rmat12_auto", M.rmat(12), -1))
    lap, _ = M.laplacian27(12)
    out.append(("lap27_12_auto", lap, -1))
    out.append(("all_sigmas_probe", M.from_row_counts(_counts(rng, 800, 0, 40), 800, 14), 19))
    # runs of thousands of empty rows inside one tile's row span (R-MAT at scale: half of all rows are empty): the
    # format kernels switch from walking the rows to binary searches beyond 2048 rows per tile
    rng2 = np.random.default_rng(77)   # own stream: the cases above keep the inputs their golden vectors were made from
    c = _counts(rng2, 12000, 1, 9)
    c[100:4100] = 0
    c[6000:9000] = 0
    c[9000] = 700
    c[-2500:] = 0
    out.append(("long_empty_runs_s6", M.from_row_counts(c, 3000, 16), 6))
    c = np.zeros(20000, np.int64)
    c[::2500] = 40
    out.append(("sparse_rows_huge_span_auto", M.from_row_counts(c, 500, 17), -1))
    return out


def sigma_sweep_case():
    rng = np.random.default_rng(99)
    c = rng.integers(0, 50, size=1200)
    c[17] = 3000
    return M.from_row_counts(c, 1200, 15, name="sigma_sweep")
