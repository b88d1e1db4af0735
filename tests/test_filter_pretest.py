"""CPU property test of the FILTER epilogue's comparison (csrc/gemm_sm100.cu, fused decode + top-K).

The kernel must keep every cell whose logit z = fl(acc + bias) is >= the playlist's threshold thr (a lower bound of its
k-th largest logit).  It does not evaluate that comparison: the scan and the queue examine fl(fl(acc + nthr) + bias) >=
-4e-7 |bias| with nthr = -(thr - (2e-6 + 4e-7 |thr|)) formed once per CTA (all fp32, round to nearest).  This test restates
both forms in NumPy float32 and checks, on random and adversarial operands (cells exactly at, one ulp above and below the
threshold, large magnitudes, infinities), that the kernel's form never drops a cell the exact form keeps, and that what
it keeps in addition lies within the stated slack below the threshold."""
import numpy as np

f32 = np.float32


def nthr_of(thr):
    thr = thr.astype(f32)
    slack = (f32(2e-6) + f32(4e-7) * np.abs(thr)).astype(f32)
    loose = -((thr - slack).astype(f32))
    return np.where(np.abs(thr) <= f32(3.0e38), loose, -thr).astype(f32)


def kernel_keeps(acc, bias, thr):
    d = (acc.astype(f32) + nthr_of(thr)).astype(f32)                     # FADD2 of the scan / __fadd_rn of the queue pass
    return (d + bias.astype(f32)).astype(f32) >= (f32(-4e-7) * np.abs(bias.astype(f32))).astype(f32)


def exact_keeps(acc, bias, thr):
    return (acc.astype(f32) + bias.astype(f32)).astype(f32) >= thr.astype(f32)


def _cases(rng, n, scale):
    bias = (rng.normal(0, scale, n)).astype(f32)
    thr = (rng.normal(0, scale, n)).astype(f32)
    # accumulators that land the logit within a few ulps of the threshold, on either side
    z = thr.copy()
    steps = rng.integers(-3, 4, n)
    for _ in range(3):
        z = np.where(steps > 0, np.nextafter(z, f32(np.inf)), np.where(steps < 0, np.nextafter(z, f32(-np.inf)), z)).astype(f32)
        steps = steps - np.sign(steps)
    acc = (z - bias).astype(f32)
    return acc, bias, thr


def test_pretest_never_drops_a_cell_the_exact_comparison_keeps():
    rng = np.random.default_rng(0)
    with np.errstate(invalid="ignore", over="ignore"):         # inf - inf below is the point of the infinity cases
        for scale in (1e-3, 0.05, 1.0, 8.0, 60.0, 4000.0, 3e5):
            acc, bias, thr = _cases(rng, 400_000, scale)
            ex, ke = exact_keeps(acc, bias, thr), kernel_keeps(acc, bias, thr)
            assert not (ex & ~ke).any(), scale
            # what is kept in addition is within the slack below the threshold (a longer list, never a wrong one)
            extra = ke & ~ex
            z = (acc + bias).astype(f32)
            tol = 4e-6 + 1.5e-6 * (np.abs(thr) + np.abs(bias))
            assert (thr[extra].astype(np.float64) - z[extra].astype(np.float64) <= tol[extra]).all(), scale
            # unrelated operands: the two forms agree except inside the slack
            acc2 = rng.normal(0, scale, len(acc)).astype(f32)
            ex2, ke2 = exact_keeps(acc2, bias, thr), kernel_keeps(acc2, bias, thr)
            assert not (ex2 & ~ke2).any(), scale
        # thresholds of rows that keep everything (-inf), nothing (+inf); rows past the item range carry a NaN bias
        acc = rng.normal(0, 3, 1000).astype(f32)
        bias = rng.normal(0, 1, 1000).astype(f32)
        assert kernel_keeps(acc, bias, np.full(1000, -np.inf, f32)).all()
        assert not kernel_keeps(acc, bias, np.full(1000, np.inf, f32)).any()
        assert not kernel_keeps(acc, np.full(1000, np.nan, f32), rng.normal(0, 1, 1000).astype(f32)).any()


def test_row_maximum_form_equals_any_cell():
    """The scan tests max_k(acc_k + nthr_k) + bias once per lane instead of every cell: rounding is monotonic, so the
    lane passes exactly when one of its 16 cells does."""
    rng = np.random.default_rng(1)
    acc = rng.normal(0, 1, (50_000, 16)).astype(f32)
    thr = (rng.normal(2.5, 0.3, (50_000, 16))).astype(f32)
    bias = rng.normal(0, 1, (50_000, 1)).astype(f32)
    d = (acc + nthr_of(thr)).astype(f32)
    lane = (d.max(1, keepdims=True) + bias).astype(f32) >= (f32(-4e-7) * np.abs(bias)).astype(f32)
    cells = kernel_keeps(acc, np.broadcast_to(bias, acc.shape), thr)
    assert np.array_equal(lane[:, 0], cells.any(1))
    assert cells.any()
