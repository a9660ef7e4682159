"""CPU oracle for the InstantSfM BA / GP Levenberg-Marquardt hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``instantsfm_b200/`` may import this package.
The only permitted users are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``, and there only as the
checker or the reported CPU baseline -- never as the thing that is shipped.

What it restates (fp64, torch-CPU / numpy / scipy, no bae, no pypose):

* residual definitions  -- /root/reference/instantsfm/utils/cost_function.py:22-177
* problem construction  -- /root/reference/instantsfm/processors/bundle_adjustment.py:66-113
                           /root/reference/instantsfm/processors/global_positioning.py:85-152
* optimiser set-up      -- bundle_adjustment.py:116-119, global_positioning.py:158-161
* outer loop / stop rule-- bundle_adjustment.py:128-141, global_positioning.py:172-183
* write-back            -- bundle_adjustment.py:18-36, global_positioning.py:41-43,199-206

PARITY UNPINNED (LM semantics).  The arithmetic of ``optimizer.step`` lives in two
third-party packages that are absent from /root/reference and cannot be installed here
(no network): ``bae`` (github.com/zitongzhan/bae, unpinned HEAD, README.md:67-70) and
``pypose`` (branch ``bae``, pyproject.toml:38).  The reference ships no test, golden
vector or fixture for this path.  What IS pinned:

* the residual functions: ``tests/golden/reference_cost_functions.npz`` was produced by
  importing the reference's own ``cost_function.py`` (with ``bae``/``pyceres`` stubbed,
  see ``tests/golden/make_reference_golden.py``) and ``oracle.camera_models`` must
  reproduce it to 1e-12;
* ``get_camera_model_info`` tables (same script);
* the oracle against itself: autograd vs finite differences, retraction consistency,
  cost monotonicity, noise-free recovery, direct-vs-PCG agreement.

What is restated from public pypose/bae semantics and therefore UNVERIFIED (assumption
ledger):

 1. ``rotate_quat(p, pose7)`` = R(q) p + t with pose7 = [t(3), q = (x, y, z, w)];
    its tangent is the left perturbation X <- Exp([dtau, dphi]) X, translation first.
 2. Unknown ordering = nn.Parameter registration order: BA [pose, points],
    GP [translations, points, scales].
 3. Damping is multiplicative on diag(J^T J), after clamp(1e-6, 1e32), and cumulative
    across rejected trials inside one ``step``.
 4. PCG = scalar-Jacobi preconditioned CG, x0 = 0, stop ||r|| < tol * ||b||.
 5. TrustRegion defaults high=0.5, low=1e-3, factor=0.5, min=1e-6.
 6. ``step`` returns the post-step robust cost sum_i rho(||r_i||^2) (no 1/2).
 7. FastTriggs: r_i <- sqrt(rho'(||r_i||^2)) r_i, J_i likewise.
 8. bae solves the full (camera + point) system, no Schur complement.
"""
