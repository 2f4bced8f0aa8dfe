"""CPU experiment for DESIGN.md section 8 item 3: can render_backward accumulate the raw moments
(sum w, sum w dx, sum w dy, sum w dx^2, sum w dx dy, sum w dy^2; w = dL_dG * G per pixel) in fp32 and apply the conic
factors once per Gaussian, instead of accumulating the reference's per-pixel products (backward.cu:585-595)?
Compares both fp32 formulations against an fp64 evaluation of the reference formulas on synthetic footprints."""
import numpy as np

rng = np.random.default_rng(0)
N = 4000
err_ref, err_mom, scale = [], [], []
f32 = np.float32
for _ in range(N):
    # random 2-D covariance: sigma 0.5..30 px, anisotropy up to 30:1, + 0.3 dilation (forward.cu:110-111)
    s1 = np.exp(rng.uniform(np.log(0.5), np.log(30.0)))
    s2 = s1 / np.exp(rng.uniform(0.0, np.log(30.0)))
    th = rng.uniform(0, np.pi)
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    cov = R @ np.diag([s1 * s1, s2 * s2]) @ R.T + 0.3 * np.eye(2)
    con = np.linalg.inv(cov)
    A, B, Cc = f32(con[0, 0]), f32(con[0, 1]), f32(con[1, 1])
    rad = int(np.ceil(3 * np.sqrt(max(np.linalg.eigvalsh(cov)))))
    rad = min(rad, 60)
    mx, my = rng.uniform(0, 1, 2)
    xs = np.arange(-rad, rad + 1)
    px, py = np.meshgrid(xs, xs)
    dx = (f32(mx) - px.astype(f32)).ravel()
    dy = (f32(my) - py.astype(f32)).ravel()
    power = f32(-0.5) * (A * dx * dx + Cc * dy * dy) - B * dx * dy
    G = np.exp(power.astype(f32)).astype(f32)
    opac = f32(rng.uniform(0.05, 1.0))
    keep = (power <= 0) & (opac * G >= 1.0 / 255.0)
    if keep.sum() < 4:
        continue
    dx, dy, G = dx[keep], dy[keep], G[keep]
    dL_dopa = rng.normal(size=G.shape).astype(f32) * f32(rng.uniform(0.1, 1.0))  # per-pixel upstream term (any sign)
    dL_dG = opac * dL_dopa
    # --- reference per-pixel products (fp32), summed in fp32 sequentially vs fp64 truth
    gdx, gdy = G * dx, G * dy
    t0 = dL_dG * (-gdx * A - gdy * B)
    t1 = dL_dG * (-gdy * Cc - gdx * B)
    t2 = f32(-0.5) * gdx * dx * dL_dG
    t3 = f32(-0.5) * gdx * dy * dL_dG
    t4 = f32(-0.5) * gdy * dy * dL_dG
    t5 = G * dL_dopa
    d = lambda a: a.astype(np.float64)
    truth = np.array([(d(dL_dG) * (-(d(G) * d(dx)) * d(A) - (d(G) * d(dy)) * d(B))).sum(),
                      (d(dL_dG) * (-(d(G) * d(dy)) * d(Cc) - (d(G) * d(dx)) * d(B))).sum(),
                      (-0.5 * d(G) * d(dx) * d(dx) * d(dL_dG)).sum(), (-0.5 * d(G) * d(dx) * d(dy) * d(dL_dG)).sum(),
                      (-0.5 * d(G) * d(dy) * d(dy) * d(dL_dG)).sum(), (d(G) * d(dL_dopa)).sum()])
    s32 = lambda a: np.cumsum(a, dtype=f32)[-1]  # sequential fp32 accumulation
    ref32 = np.array([s32(t0), s32(t1), s32(t2), s32(t3), s32(t4), s32(t5)], dtype=np.float64)
    # --- moments in fp32, conic factors applied once
    w = dL_dopa * G  # = t5 terms; dL_dG * G = opac * w
    S0, S1, S2 = s32(w), s32(w * dx), s32(w * dy)
    S3, S4, S5 = s32(w * dx * dx), s32(w * dx * dy), s32(w * dy * dy)
    mom = np.array([-(opac * (A * S1 + B * S2)), -(opac * (Cc * S2 + B * S1)), f32(-0.5) * opac * S3, f32(-0.5) * opac * S4,
                    f32(-0.5) * opac * S5, S0], dtype=np.float64)
    err_ref.append(ref32 - truth)
    err_mom.append(mom - truth)
    scale.append(truth)
err_ref, err_mom, scale = map(np.array, (err_ref, err_mom, scale))
names = ["dmean2D.x", "dmean2D.y", "dconic.x", "dconic.y", "dconic.z", "dopacity"]
print(f"{len(scale)} footprints; relative L2 error over all Gaussians (vs fp64):")
for k, n in enumerate(names):
    den = np.linalg.norm(scale[:, k])
    print(f"  {n:10s} reference-style fp32 {np.linalg.norm(err_ref[:, k]) / den:.2e}   moments fp32 {np.linalg.norm(err_mom[:, k]) / den:.2e}"
          f"   worst single Gaussian (moments, rel. to its own value) {np.max(np.abs(err_mom[:, k]) / (np.abs(scale[:, k]) + 1e-30)):.2e}")
