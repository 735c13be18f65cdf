// gtf_math.cuh -- per-edge / per-pair fp64 algebra of the message-passing hot path.
//
// Pure functions, usable from device code (kernels in gtf_kernels.cu) and -- for CPU-side unit tests of
// the algebra against the oracle -- from host code (csrc/gtf_hostmath.cpp compiles this header with g++).
// Everything is written for the covariance structure the reference actually stores:
//     [ p00 p01  0  ]
//     [ p01 p11  0  ]      (row/col 2 zeroed: helper.py:423-425, extrapolate_merged_states.py:363-365)
//     [  0   0  p22 ]
// so 3x3 inverses collapse to a closed-form 2x2 inverse plus a reciprocal, and sin/cos(atan2(s, c)) are
// taken algebraically (s/h, c/h).  Reference line numbers are relative to /root/reference/src.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define GTF_HD __host__ __device__ __forceinline__
#else
#define GTF_HD inline
#endif

struct GtfGeom {
    double sigma0xy, sigma0rz, sigma0rz2, endcap;
};

struct GtfState {          // one Gaussian component: parabola (a,b,c), slope tau, block covariance
    double a, b, c, tau, p00, p01, p11, p22;
};

GTF_HD double gtf_sq(double v) { return v * v; }
#ifdef __CUDA_ARCH__
#define GTF_RSQRT(x) rsqrt(x)
#else
#define GTF_RSQRT(x) (1.0 / sqrt(x))
#endif

// Highland multiple-scattering variance of the track direction (helper.py:402-415,
// extrapolate_merged_states.py:114-124, extract_track_candidates.py:244-255).
// dr, dz: segment; xk: global x of the far hit; endcap_side_z: the z whose |z| >= endcap selects tan(theta).
GTF_HD double gtf_var_ms(double a, double b, double xk, double dr, double dz, double endcap_side_z, double endcap)
{
    double hyp = sqrt(dr * dr + dz * dz);
    double sin_t = fabs(dr) / hyp;
    double kb = 2.0 * a * xk + b;
    double t = 1.0 + kb * kb;
    double rt = GTF_RSQRT(t);
    double kappa = (2.0 * a) * (rt * rt * rt);                                // 2a / t^(3/2)
    double q = (13.6 * 1e-3 * 0.1414213562373095048801688724 / 0.3) * kappa;  // sqrt(0.02); constant folded
    double v = sin_t * (q * q);
    if (fabs(endcap_side_z) >= endcap) v = v * (fabs(dr) / fabs(dz));
    return v;
}

// the same with the per-edge geometry (sin(theta) = |dr| / hypot(dr, dz) and the end-cap ratio |dr| / |dz|) computed
// once per edge instead of per message: same operations in the same order as gtf_var_ms
GTF_HD void gtf_var_ms_geo(double dr, double dz, double &sin_t, double &rdz)
{
    double hyp = sqrt(dr * dr + dz * dz);
    sin_t = fabs(dr) / hyp;
    rdz = fabs(dr) / fabs(dz);
}
GTF_HD double gtf_var_ms_pre(double a, double b, double xk, double sin_t, double rdz, double endcap_side_z, double endcap)
{
    double kb = 2.0 * a * xk + b;
    double t = 1.0 + kb * kb;
    double rt = GTF_RSQRT(t);
    double kappa = (2.0 * a) * (rt * rt * rt);                                // 2a / t^(3/2)
    double q = (13.6 * 1e-3 * 0.1414213562373095048801688724 / 0.3) * kappa;  // sqrt(0.02); constant folded
    double v = sin_t * (q * q);
    if (fabs(endcap_side_z) >= endcap) v = v * rdz;
    return v;
}

// variance of tau = dz/dr from the four measurement errors (helper.py:317-330, extrapolate...py:344-358)
GTF_HD double gtf_var_tau(double dz, double dr, double z_node, double z_nb, const GtfGeom &g)
{
    double sr = g.sigma0rz, sz = g.sigma0rz2, srn = g.sigma0rz, szn = g.sigma0rz2;
    if (fabs(z_node) >= g.endcap) { sz = g.sigma0rz; sr = g.sigma0rz2; }
    if (fabs(z_nb) >= g.endcap) { szn = g.sigma0rz; srn = g.sigma0rz2; }
    double j1 = 1.0 / dr, j3 = dz / (dr * dr);
    return j1 * j1 * (sz * sz) + j1 * j1 * (szn * szn) + j3 * j3 * (sr * sr) + j3 * j3 * (srn * srn);
}

// ------------------------------------------------------------------------------------------------
// Seeding: helper.py:354-441 for one (node, key) pair.  `tau`/`var_tau_sq` come from the *mirrored*
// neighbour (quirk 5: key N[d-1-i] gets the tau of N[i]) and are supplied by the caller.
GTF_HD void gtf_seed_entry(double xA, double yA, double zA, double rA, double xk, double yk, double zk, double rk,
                           double tau, double var_tau_sq, const GtfGeom &g, GtfState &o)
{
    double rho = sqrt(xA * xA + yA * yA);
    double ca = rho > 0.0 ? xA / rho : 1.0, sa = rho > 0.0 ? yA / rho : 0.0;
    double x0 = (0.0 - xA) * ca + (0.0 - yA) * sa;
    double xB = (xk - xA) * ca + (yk - yA) * sa;
    double mB = -(xk - xA) * sa + (yk - yA) * ca;
    // H = [[x0^2/2, x0, 1], [0, 0, 1], [xB^2/2, xB, 1]]; measurement (0, 0, mB)  ->  closed-form inverse
    double det = 0.5 * x0 * xB * (x0 - xB);          // det of the 2x2 system after subtracting row 1
    // rows of H^-1:  Hi0 = ( xB, -(xB - x0), -x0) / det ;  Hi1 = (-xB^2/2, (xB^2 - x0^2)/2, x0^2/2) / det ; Hi2 = (0,1,0)
    double h00 = xB / det, h01 = (x0 - xB) / det, h02 = -x0 / det;
    double h10 = -0.5 * xB * xB / det, h11 = 0.5 * (xB * xB - x0 * x0) / det, h12 = 0.5 * x0 * x0 / det;
    o.a = h02 * mB;
    o.b = h12 * mB;
    o.c = 0.0 * mB;                                   // Hi2 . (0, 0, mB) = 0 (NaN-propagating like the reference)
    double sO = 16.0, sA = g.sigma0xy * g.sigma0xy;   // sigmaO = 4 mm (helper.py:243)
    double var_ms = gtf_var_ms(o.a, o.b, xk, rA - rk, zA - zk, zA, g.endcap);
    o.p00 = h00 * h00 * sO + h01 * h01 * sA + h02 * h02 * sA;
    o.p01 = h00 * h10 * sO + h01 * h11 * sA + h02 * h12 * sA;
    o.p11 = h10 * h10 * sO + h11 * h11 * sA + h12 * h12 * sA + var_ms;
    o.tau = tau;
    o.p22 = var_tau_sq + var_ms;
}

// ------------------------------------------------------------------------------------------------
// Parabolic-model seeding of the KL look-up-table training pipeline, one (node, neighbour) pair:
// learn_KL_parabolic_model/src/generate_training_data/utils.py:221-299 (compute_track_state_estimates) with :197-218
// (rotate_track: the origin -> node edge becomes the x axis) -- NOT the seeding of the reconstruction (gtf_seed_entry):
// H rows are (x^2, x, 1) without the 1/2, the measurement errors are S = diag(sigma0^2, sigmaA^2, sigmaB^2) and the
// covariance H^-1 S H^-T is a FULL 3x3 matrix (row-major cov9).
GTF_HD void gtf_seed_parabolic(double xn, double yn, double xb, double yb, double sigma0, double sigmaA, double sigmaB,
                               double *sv3, double *cov9)
{
    const double two_pi = 2.0 * 3.14159265358979323846;
    double ang = atan2(yn - 0.0, xn - 0.0);                       // :191-194 getAngleBetweenPoints((0, 0), node)
    while (ang < 0.0) ang += two_pi;                              // :184-187 angle_trunc
    const double angle = two_pi - ang;                            // :209
    const double ca = cos(angle), sa = sin(angle);
    const double xr_n = xn * ca - yn * sa, yr_n = xn * sa + yn * ca;   // :214-216
    const double xr_b = xb * ca - yb * sa, yr_b = xb * sa + yb * ca;
    const double xr_0 = 0.0 * ca - 0.0 * sa;
    const double x0 = xr_0 - xr_n, xB = xr_b - xr_n, mB = yr_b - yr_n; // :262-269 translate: the node becomes the origin
    // H = [[x0^2, x0, 1], [0, 0, 1], [xB^2, xB, 1]] (:281-283), closed-form inverse
    const double iD = 1.0 / (x0 * xB * (x0 - xB));
    const double h[3][3] = {{xB * iD, (x0 - xB) * iD, -x0 * iD},
                            {-(xB * xB) * iD, (xB * xB - x0 * x0) * iD, (x0 * x0) * iD},
                            {0.0, 1.0, 0.0}};
    const double m[3] = {0.0, 0.0, mB};                           // :280 [m_0, m_A, m_B]
    const double S[3] = {sigma0 * sigma0, sigmaA * sigmaA, sigmaB * sigmaB};
    for (int i = 0; i < 3; i++) {
        sv3[i] = h[i][0] * m[0] + h[i][1] * m[1] + h[i][2] * m[2];     // :287
        for (int j = 0; j < 3; j++)
            cov9[3 * i + j] = h[i][0] * S[0] * h[j][0] + h[i][1] * S[1] * h[j][1] + h[i][2] * S[2] * h[j][2];   // :288
    }
}

// ------------------------------------------------------------------------------------------------
// Extrapolate + chi2 gate + Kalman update for one edge u -> v (extrapolate_merged_states.py:26-402).
struct GtfExtrapOut {
    double chi2, lik, var_ms;
    GtfState s;        // updated state stored at the receiver
    int pass;
};

// Split in two so a kernel can issue the next edge's loads between the halves: gtf_extrap_jac (geometry, Jacobian F,
// once-propagated state) and gtf_extrap_update (covariance propagation, gate, filterpy predict + update, tau).
struct GtfJac {
    double f00, f01, f02, f10, f11, f12, f20, f21, f22;
    double xe0, xe1, xe2;
};
GTF_HD void gtf_extrap_jac(double ux, double uy, double vx, double vy, double a, double b, double c, GtfJac &J)
{
    // rotation into the source frame (:41,52) -- only x_A is live; 1/rho and 1/h as reciprocal square roots
    double rho2 = ux * ux + uy * uy;
    double irho = GTF_RSQRT(rho2);
    double ca = rho2 > 0.0 ? ux * irho : 1.0, sa = rho2 > 0.0 ? uy * irho : 0.0;
    double xA = (vx - ux) * ca + (vy - uy) * sa;
    // phi between the two radius vectors (:59): sin/cos taken algebraically
    double cr = ux * vy - uy * vx, dt = ux * vx + uy * vy;
    double h2 = cr * cr + dt * dt;
    double ih = GTF_RSQRT(h2);
    double sp = h2 > 0.0 ? cr * ih : 0.0, cp = h2 > 0.0 ? dt * ih : 1.0;
    double xp = xA + c * sp, Vx = cp + b * sp, Ax = a * sp;                     // :63-65
    double Vx2 = Vx * Vx;
    double iV = 1.0 / Vx, iV2 = iV * iV;
    double s_star = (-xp * (2.0 * Vx2 + Ax * xp)) * (0.5 * iV2 * iV);           // :68 (one reciprocal of Vx serves all)
    // ds*/d(a,b,c) (:82-86); numer == xp, denom == Vx
    double ds_da = -(sp * xp * xp) * iV2 * iV;
    double ds_db = (sp * xp) * (1.0 + 3.0 * a * sp * xp * iV2) * iV2;
    double ds_dc = -sp * (1.0 + 2.0 * a * sp * xp * iV2) * iV;
    // da'/d. (:89-92)
    double den = cp + (2.0 * a + b) * sp, id = 1.0 / den, id2 = id * id, id4 = id2 * id2;
    J.f00 = (id2 * id) * (1.0 - (6.0 * a * sp) * (s_star + a * ds_da) * id);
    J.f01 = (-3.0 * a * sp * (2.0 * a * ds_db + 1.0)) * id4;
    J.f02 = (-6.0 * sp * ds_dc * a * a) * id4;
    // db'/d. (:95-99)
    double w = 2.0 * a * s_star + b;
    den = cp + w * sp;
    id = 1.0 / den;
    double br = cp - (sp * (-sp + w * cp)) * id;
    J.f10 = (2.0 * (s_star + a * ds_da) * br) * id;
    J.f11 = ((1.0 + 2.0 * a * ds_da) * br) * id;
    J.f12 = (2.0 * a * ds_dc * br) * id;
    // dc'/d. (:102-105)
    br = cp * (2.0 * a + b) - sp;
    J.f20 = ds_da * br + s_star * s_star * cp;
    J.f21 = ds_db * br + s_star * cp;
    J.f22 = ds_dc * br + cp;
    // first propagation of the state (:129): x_e = F x
    J.xe0 = J.f00 * a + J.f01 * b + J.f02 * c;
    J.xe1 = J.f10 * a + J.f11 * b + J.f12 * c;
    J.xe2 = J.f20 * a + J.f21 * b + J.f22 * c;
}
GTF_HD void gtf_extrap_update(const GtfJac &J, double dr, double dz, double uz, double vz, double p00, double p01,
                              double p11_eff, double p22, double var_ms, double chi2_cut, const GtfGeom &g, GtfExtrapOut &o)
{
    const double f00 = J.f00, f01 = J.f01, f02 = J.f02, f10 = J.f10, f11 = J.f11, f12 = J.f12, f20 = J.f20, f21 = J.f21,
                 f22 = J.f22, xe0 = J.xe0, xe1 = J.xe1, xe2 = J.xe2;
    // P_e = F P F^T with block P (p11 already carries the accumulated multiple-scattering term, :127-130)
#define GTF_FPFT(ri0, ri1, ri2, rj0, rj1, rj2) \
    ((ri0) * (p00 * (rj0) + p01 * (rj1)) + (ri1) * (p01 * (rj0) + p11_eff * (rj1)) + (ri2) * p22 * (rj2))
    double e00 = GTF_FPFT(f00, f01, f02, f00, f01, f02);
    double e01 = GTF_FPFT(f00, f01, f02, f10, f11, f12);
    double e02 = GTF_FPFT(f00, f01, f02, f20, f21, f22);
    double e11 = GTF_FPFT(f10, f11, f12, f10, f11, f12);
    double e12 = GTF_FPFT(f10, f11, f12, f20, f21, f22);
    double e22 = GTF_FPFT(f20, f21, f22, f20, f21, f22);
#undef GTF_FPFT
    double R = g.sigma0xy * g.sigma0xy;
    double res = 0.0 - xe2;                                                     // :137
    double S = e22 + R;                                                         // :138
    double chi2 = (res / S) * res;                                              // :140
    o.chi2 = chi2;
    o.var_ms = var_ms;
    o.pass = chi2 <= chi2_cut;                                                  // :298 (NaN -> fail)
    if (!o.pass) return;
    o.lik = exp(-0.5 * chi2) * GTF_RSQRT(2.0 * 3.14159265358979323846 * fabs(S)); // :302-304 (1/sqrt as one op)
    // filterpy predict(): F applied a second time, + Q = diag(0, var_ms, 0)    (:321)
    double x0 = f00 * xe0 + f01 * xe1 + f02 * xe2;
    double x1 = f10 * xe0 + f11 * xe1 + f12 * xe2;
    double x2 = f20 * xe0 + f21 * xe1 + f22 * xe2;
    // G = F * P_e (rows), then P' = G F^T ; P_e symmetric
#define GTF_G(r0, r1, r2, c0, c1, c2) ((r0) * (c0) + (r1) * (c1) + (r2) * (c2))
    double g00 = GTF_G(f00, f01, f02, e00, e01, e02), g01 = GTF_G(f00, f01, f02, e01, e11, e12),
           g02 = GTF_G(f00, f01, f02, e02, e12, e22);
    double g10 = GTF_G(f10, f11, f12, e00, e01, e02), g11 = GTF_G(f10, f11, f12, e01, e11, e12),
           g12 = GTF_G(f10, f11, f12, e02, e12, e22);
    double g20 = GTF_G(f20, f21, f22, e00, e01, e02), g21 = GTF_G(f20, f21, f22, e01, e11, e12),
           g22 = GTF_G(f20, f21, f22, e02, e12, e22);
    double q00 = GTF_G(g00, g01, g02, f00, f01, f02);
    double q01 = GTF_G(g00, g01, g02, f10, f11, f12);
    double q02 = GTF_G(g00, g01, g02, f20, f21, f22);
    double q11 = GTF_G(g10, g11, g12, f10, f11, f12) + var_ms;
    double q12 = GTF_G(g10, g11, g12, f20, f21, f22);
    double q22 = GTF_G(g20, g21, g22, f20, f21, f22);
#undef GTF_G
    // filterpy update(z = 0), H = [0 0 1], Joseph form (:322)
    double y = 0.0 - x2;
    double SI = 1.0 / (q22 + R);
    double K0 = q02 * SI, K1 = q12 * SI, K2 = q22 * SI;
    o.s.a = x0 + K0 * y;
    o.s.b = x1 + K1 * y;
    o.s.c = x2 + K2 * y;
    double t00 = q00 - K0 * q02, t01 = q01 - K0 * q12, t02 = q02 - K0 * q22;
    double t11 = q11 - K1 * q12, t12 = q12 - K1 * q22;
    o.s.p00 = t00 - K0 * t02 + K0 * R * K0;
    o.s.p01 = t01 - K1 * t02 + K0 * R * K1;
    o.s.p11 = t11 - K1 * t12 + K1 * R * K1;
    // tau and its variance (:326-365)
    // one reciprocal of dr serves tau and both Jacobian terms of var(tau) (each within 1 ulp of the quotient)
    {
        double sr = g.sigma0rz, sz = g.sigma0rz2, srn = g.sigma0rz, szn = g.sigma0rz2;
        if (fabs(uz) >= g.endcap) { sz = g.sigma0rz; sr = g.sigma0rz2; }
        if (fabs(vz) >= g.endcap) { szn = g.sigma0rz; srn = g.sigma0rz2; }
        double j1 = 1.0 / dr, tau = dz * j1, j3 = tau * j1;
        o.s.tau = tau;
        o.s.p22 = (j1 * j1 * (sz * sz) + j1 * j1 * (szn * szn) + j3 * j3 * (sr * sr) + j3 * j3 * (srn * srn)) + var_ms;
    }
}
GTF_HD void gtf_extrapolate(double ux, double uy, double uz, double ur, double vx, double vy, double vz, double vr,
                            double a, double b, double c, double p00, double p01, double p11_eff, double p22,
                            double var_ms, double chi2_cut, const GtfGeom &g, GtfExtrapOut &o)
{
    GtfJac J;
    gtf_extrap_jac(ux, uy, vx, vy, a, b, c, J);
    gtf_extrap_update(J, vr - ur, vz - uz, uz, vz, p00, p01, p11_eff, p22, var_ms, chi2_cut, g, o);
}

// ------------------------------------------------------------------------------------------------
// clustering.py:11-78 pairwise chi2 of components i ("neighbour1", b) and j ("neighbour2", c) at node a.
// The tau part only needs per-component quantities (1/dr, tau, its measurement-error terms), so they are computed
// once per component (gtf_pair_geo) and a pair costs two divisions instead of four.
struct GtfPairGeo {
    double I, T, Q, A, C; // 1/(r_e - r_a);  tau_e = (z_e - z_a) I;  tau_e I;  I^2 sigma_z,e^2;  Q^2 sigma_r,e^2
};
GTF_HD void gtf_pair_geo(double xe, double ze, double re, double za, double ra, const GtfGeom &g, GtfPairGeo &G)
{
    double sz = g.sigma0rz2, sr = g.sigma0rz;
    if (fabs(xe) >= g.endcap) { sz = g.sigma0rz; sr = g.sigma0rz2; }     // abs(x), not abs(z): clustering.py:49-57
    G.I = 1.0 / (re - ra);
    G.T = (ze - za) * G.I;
    G.Q = G.T * G.I;
    G.A = G.I * G.I * sz * sz;
    G.C = G.Q * G.Q * sr * sr;
}
GTF_HD double gtf_pair_chi2_pre(const GtfState &si, const GtfState &sj, double xa, const GtfPairGeo &Gb, const GtfPairGeo &Gc,
                                const GtfGeom &g)
{
    double r0 = si.a - sj.a, r1 = si.b - sj.b;
    double c00 = si.p00 + sj.p00, c01 = si.p01 + sj.p01, c11 = si.p11 + sj.p11;
    double det = c00 * c11 - c01 * c01;
    double num = r0 * r0 * c11 - 2.0 * r0 * r1 * c01 + r1 * r1 * c00;
    double sza = g.sigma0rz2, sra = g.sigma0rz;
    if (fabs(xa) >= g.endcap) { sza = g.sigma0rz; sra = g.sigma0rz2; }
    double j1 = Gc.I - Gb.I, j4 = Gb.Q - Gc.Q;                            // d tau_b - tau_c / d(z_a, r_a)
    double cdt = j1 * j1 * sza * sza + Gb.A + Gc.A + j4 * j4 * sra * sra + Gb.C + Gc.C;
    double dtau = Gb.T - Gc.T;
    // num / det + dtau^2 / cdt over one common denominator: one division per pair
    return (num * cdt + dtau * dtau * det) / (det * cdt);
}
GTF_HD double gtf_pair_chi2(const GtfState &si, const GtfState &sj, double xa, double za, double ra, double xb,
                            double zb, double rb, double xc, double zc, double rc, const GtfGeom &g)
{
    GtfPairGeo Gb, Gc;
    gtf_pair_geo(xb, zb, rb, za, ra, g, Gb);
    gtf_pair_geo(xc, zc, rc, za, ra, g, Gc);
    return gtf_pair_chi2_pre(si, sj, xa, Gb, Gc, g);
}

// inverse of the 2x2 block: returns (i00, i01, i11)
GTF_HD void gtf_inv2(double p00, double p01, double p11, double &i00, double &i01, double &i11)
{
    double id = 1.0 / (p00 * p11 - p01 * p01);
    i00 = p11 * id;
    i01 = -p01 * id;
    i11 = p00 * id;
}

// clustering.py:97-105 merge_states for BOTH chains at once: (a,b,c) and (a,b,tau) share the covariance
// (quirk 4: edge_covariance is joint_vector_covariance), so c and tau are both fused with weight 1/p22.
GTF_HD void gtf_merge(const GtfState &s1, const GtfState &s2, GtfState &m)
{
    double a00, a01, a11, b00, b01, b11;
    gtf_inv2(s1.p00, s1.p01, s1.p11, a00, a01, a11);
    gtf_inv2(s2.p00, s2.p01, s2.p11, b00, b01, b11);
    double s00 = a00 + b00, s01 = a01 + b01, s11 = a11 + b11;
    double m00, m01, m11;
    gtf_inv2(s00, s01, s11, m00, m01, m11);
    double v0 = (a00 * s1.a + a01 * s1.b) + (b00 * s2.a + b01 * s2.b);
    double v1 = (a01 * s1.a + a11 * s1.b) + (b01 * s2.a + b11 * s2.b);
    double q1 = 1.0 / s1.p22, q2 = 1.0 / s2.p22;
    double mq = 1.0 / (q1 + q2);
    GtfState r;
    r.a = m00 * v0 + m01 * v1;
    r.b = m01 * v0 + m11 * v1;
    r.c = mq * (q1 * s1.c + q2 * s2.c);
    r.tau = mq * (q1 * s1.tau + q2 * s2.tau);
    r.p00 = m00;
    r.p01 = m01;
    r.p11 = m11;
    r.p22 = mq;
    m = r;
}

// clustering.py:90-94 KLDistance on the joint vector (a, b, tau).  The "trace" is of an ELEMENT-WISE
// product (numpy `*`), i.e. sum_k (c1_kk - c2_kk) * (inv2_kk - inv1_kk): reproduced, pinned by the golden CSV.
GTF_HD double gtf_kl(const GtfState &s1, const GtfState &s2)
{
    double a00, a01, a11, b00, b01, b11;
    gtf_inv2(s1.p00, s1.p01, s1.p11, a00, a01, a11);
    gtf_inv2(s2.p00, s2.p01, s2.p11, b00, b01, b11);
    double q1 = 1.0 / s1.p22, q2 = 1.0 / s2.p22;
    double tr = (s1.p00 - s2.p00) * (b00 - a00) + (s1.p11 - s2.p11) * (b11 - a11) + (s1.p22 - s2.p22) * (q2 - q1);
    double d0 = s1.a - s2.a, d1 = s1.b - s2.b, d2 = s1.tau - s2.tau;
    double s00 = a00 + b00, s01 = a01 + b01, s11 = a11 + b11;
    return tr + (d0 * d0 * s00 + 2.0 * d0 * d1 * s01 + d1 * d1 * s11 + d2 * d2 * (q1 + q2));
}

// KLDistance for GENERAL 3x3 covariances (the LUT training-data generator seeds components with full matrices:
// learn_KL_parabolic_model/.../utils.py:221-299, compute_KL_distance.py:11-21): closed-form cofactor inverse; the
// "trace" is again of the element-wise product (clustering.py:93).  Pinned by the reference's shipped golden CSV.
GTF_HD void gtf_inv3_general(const double *c, double *o)
{
    double c00 = c[4] * c[8] - c[5] * c[7], c01 = c[5] * c[6] - c[3] * c[8], c02 = c[3] * c[7] - c[4] * c[6];
    double id = 1.0 / (c[0] * c00 + c[1] * c01 + c[2] * c02);
    o[0] = c00 * id; o[1] = (c[2] * c[7] - c[1] * c[8]) * id; o[2] = (c[1] * c[5] - c[2] * c[4]) * id;
    o[3] = c01 * id; o[4] = (c[0] * c[8] - c[2] * c[6]) * id; o[5] = (c[2] * c[3] - c[0] * c[5]) * id;
    o[6] = c02 * id; o[7] = (c[1] * c[6] - c[0] * c[7]) * id; o[8] = (c[0] * c[4] - c[1] * c[3]) * id;
}
GTF_HD double gtf_kl_general(const double *m1, const double *c1, const double *m2, const double *c2)
{
    double i1[9], i2[9];
    gtf_inv3_general(c1, i1);
    gtf_inv3_general(c2, i2);
    double tr = 0.0;
    for (int k = 0; k < 3; k++) tr += (c1[4 * k] - c2[4 * k]) * (i2[4 * k] - i1[4 * k]);
    double d0 = m1[0] - m2[0], d1 = m1[1] - m2[1], d2 = m1[2] - m2[2];
    double t0 = d0 * (i1[0] + i2[0]) + d1 * (i1[3] + i2[3]) + d2 * (i1[6] + i2[6]);   // row vector times matrix first
    double t1 = d0 * (i1[1] + i2[1]) + d1 * (i1[4] + i2[4]) + d2 * (i1[7] + i2[7]);
    double t2 = d0 * (i1[2] + i2[2]) + d1 * (i1[5] + i2[5]) + d2 * (i1[8] + i2[8]);
    return tr + (t0 * d0 + t1 * d1 + t2 * d2);
}

// clustering.py:97-105 merge_states for GENERAL 3x3 covariances (the stand-alone helper; the kernels use the block form)
GTF_HD void gtf_merge_general(const double *m1, const double *c1, const double *m2, const double *c2, double *mm, double *mc)
{
    double i1[9], i2[9], sum[9];
    gtf_inv3_general(c1, i1);
    gtf_inv3_general(c2, i2);
    for (int k = 0; k < 9; k++) sum[k] = i1[k] + i2[k];
    gtf_inv3_general(sum, mc);
    double v[3];
    for (int r = 0; r < 3; r++)
        v[r] = (i1[3 * r] * m1[0] + i1[3 * r + 1] * m1[1] + i1[3 * r + 2] * m1[2]) +
               (i2[3 * r] * m2[0] + i2[3 * r + 1] * m2[1] + i2[3 * r + 2] * m2[2]);
    for (int r = 0; r < 3; r++) mm[r] = mc[3 * r] * v[0] + mc[3 * r + 1] * v[1] + mc[3 * r + 2] * v[2];
}

// ------------------------------------------------------------------------------------------------
// Candidate quality gate (extract/extract_track_candidates.py:172-328).

// regularised upper incomplete gamma Q(a, x): power series for x < a + 1, modified-Lentz continued
// fraction otherwise.  chi2.sf(x, k) = Q(k/2, x/2) (extract...py:321,325 call scipy's chi2.sf).
GTF_HD double gtf_gammq(double a, double x)
{
    if (x != x || a != a) return NAN;
    if (x <= 0.0) return 1.0;
    if (isinf(x)) return 0.0;
    double lead = exp(-x + a * log(x) - lgamma(a));
    if (x < a + 1.0) {
        double ap = a, term = 1.0 / a, sum = term;
        for (int it = 0; it < 2000; it++) {
            ap += 1.0;
            term *= x / ap;
            sum += term;
            if (fabs(term) < fabs(sum) * 1e-17) break;
        }
        return 1.0 - sum * lead;
    }
    const double tiny = 1e-300;
    double bb = x + 1.0 - a, cc = 1.0 / tiny, dd = 1.0 / bb, hh = dd;
    for (int it = 1; it < 2000; it++) {
        double an = -it * (it - a);
        bb += 2.0;
        dd = an * dd + bb;
        if (fabs(dd) < tiny) dd = tiny;
        cc = bb + an / cc;
        if (fabs(cc) < tiny) cc = tiny;
        dd = 1.0 / dd;
        double del = dd * cc;
        hh *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return lead * hh;
}

// rotate_track (:172-193) + KF_track_fit_moliere (:209-328).  co[k] = (x, y, z, r) sorted by r, largest
// first; rotated in place.  xy filter: 3-state OU model; zr filter: 2-state with the scalar process noise
// broadcast onto all four covariance entries (filterpy semantics of `g.Q = var_ms`).
GTF_HD void gtf_track_fit(double (*co)[4], int n, double sigma0xy, double sigma0rz, double endcap, double sep3d,
                          double &pval_xy, double &pval_zr)
{
    {
        const double *p1 = co[n - 1], *p2 = co[n - 2];
        double d3 = sqrt(gtf_sq(p1[0] - p2[0]) + gtf_sq(p1[1] - p2[1]) + gtf_sq(p1[2] - p2[2]));
        if (d3 < sep3d) p2 = co[n - 3];
        double dxy = sqrt(gtf_sq(p2[0] - p1[0]) + gtf_sq(p2[1] - p1[1]));
        double cxy = dxy > 0.0 ? (p2[0] - p1[0]) / dxy : 1.0, sxy = dxy > 0.0 ? (p2[1] - p1[1]) / dxy : 0.0;
        double dzr = sqrt(gtf_sq(p2[2] - p1[2]) + gtf_sq(p2[3] - p1[3]));
        double czr = dzr > 0.0 ? (p2[3] - p1[3]) / dzr : 1.0, szr = dzr > 0.0 ? (p2[2] - p1[2]) / dzr : 0.0;
        for (int k = 0; k < n; k++) {
            double x = co[k][0], y = co[k][1], z = co[k][2], r = co[k][3];
            co[k][0] = x * cxy + y * sxy;
            co[k][1] = -x * sxy + y * cxy;
            co[k][3] = r * czr + r * szr;   // :190 (r appears twice in the reference)
            co[k][2] = -z * szr + z * czr;  // :191
        }
    }
    const double Rxy = sigma0xy * sigma0xy, Rzr = sigma0rz * sigma0rz;
    double x0 = co[0][1], x1 = 0.0, x2 = 0.0;                               // f.x
    double P00 = Rxy, P01 = 0, P02 = 0, P11 = 1, P12 = 0, P22 = 1;          // f.P (symmetric)
    double g0 = co[0][3], g1 = 0.0;                                         // g.x
    double G00 = Rzr, G01 = 0.0, G10 = 0.0, G11 = 1000.0;                   // g.P (not kept symmetric by the scalar Q)
    double chi_xy = 0.0, chi_zr = 0.0;
    for (int i = 0; i + 1 < n; i++) {
        double xa = co[i][0], ya = co[i][1], xb = co[i + 1][0], yb = co[i + 1][1];
        // parabola through the origin and the two hits (:202-204 with x1 = y1 = 0)
        double den = (-xa) * (-xb) * (xa - xb);
        double pa = (xb * ya - xa * yb) / den;
        double pb = (-(xb * xb) * ya + (xa * xa) * yb) / den;
        double dr = co[i + 1][3] - co[i][3], dz = co[i + 1][2] - co[i][2];
        double var_ms = gtf_var_ms(pa, pb, xb, dr, dz, co[i + 1][2], endcap);
        double dx = xb - xa, alpha = 0.1;
        double e1 = exp(-fabs(dx) * alpha), f1 = (1.0 - e1) / alpha, g1f = (fabs(dx) - f1) / alpha;
        double sw2 = 1e-5 * 1e-5, dx2 = dx * dx, dxw2 = dx2 * sw2;
        double Q02 = 0.5 * dxw2, Q01 = dx * (var_ms + Q02), Q12 = dx * sw2;
        double Q00 = dx2 * (var_ms + 0.25 * dxw2), Q11 = var_ms + dxw2, Q22 = sw2;
        // predict: F = [[1, dx, g1f], [0, 1, f1], [0, 0, e1]]
        double y0 = x0 + dx * x1 + g1f * x2, y1 = x1 + f1 * x2, y2 = e1 * x2;
        // A = F P
        double A00 = P00 + dx * P01 + g1f * P02, A01 = P01 + dx * P11 + g1f * P12, A02 = P02 + dx * P12 + g1f * P22;
        double A10 = P01 + f1 * P02, A11 = P11 + f1 * P12, A12 = P12 + f1 * P22;
        double A20 = e1 * P02, A21 = e1 * P12, A22 = e1 * P22;
        (void)A20; (void)A21; (void)A10;
        double N00 = A00 + dx * A01 + g1f * A02 + Q00, N01 = A01 + f1 * A02 + Q01, N02 = e1 * A02 + Q02;
        double N11 = A11 + f1 * A12 + Q11, N12 = e1 * A12 + Q12, N22 = e1 * A22 + Q22;
        // update with H = [1 0 0]
        double meas = yb, yres = meas - y0;
        double SI = 1.0 / (N00 + Rxy);
        double K0 = N00 * SI, K1 = N01 * SI, K2 = N02 * SI;
        x0 = y0 + K0 * yres; x1 = y1 + K1 * yres; x2 = y2 + K2 * yres;
        // Joseph form with I-KH = [[1-K0,0,0],[-K1,1,0],[-K2,0,1]]
        double c0 = 1.0 - K0;
        double T00 = c0 * N00, T01 = c0 * N01, T02 = c0 * N02;
        double T10 = N01 - K1 * N00, T11 = N11 - K1 * N01, T12 = N12 - K1 * N02;
        double T20 = N02 - K2 * N00, T21 = N12 - K2 * N01, T22 = N22 - K2 * N02;
        P00 = T00 * c0 + K0 * Rxy * K0;
        P01 = T00 * (-K1) + T01 + K0 * Rxy * K1;
        P02 = T00 * (-K2) + T02 + K0 * Rxy * K2;
        P11 = T10 * (-K1) + T11 + K1 * Rxy * K1;
        P12 = T10 * (-K2) + T12 + K1 * Rxy * K2;
        P22 = T20 * (-K2) + T22 + K2 * Rxy * K2;
        (void)T21;
        double res = meas - x0;                                  // post-fit residual (:292-296)
        chi_xy += (res / (P00 + Rxy)) * res;
        // zr filter: F = [[1, dz], [0, 1]], Q scalar -> + var_ms on every entry (:299-316)
        double h0 = g0 + dz * g1, h1 = g1;
        double B00 = G00 + dz * G10, B01 = G01 + dz * G11, B10 = G10, B11 = G11;
        double M00 = B00 + B01 * dz + var_ms, M01 = B01 + var_ms, M10 = B10 + B11 * dz + var_ms, M11 = B11 + var_ms;
        double gm = co[i + 1][3], gy = gm - h0;
        double gSI = 1.0 / (M00 + Rzr);
        double L0 = M00 * gSI, L1 = M10 * gSI;
        g0 = h0 + L0 * gy; g1 = h1 + L1 * gy;
        double d0 = 1.0 - L0;
        double U00 = d0 * M00, U01 = d0 * M01, U10 = M10 - L1 * M00, U11 = M11 - L1 * M01;
        G00 = U00 * d0 + L0 * Rzr * L0;
        G01 = U00 * (-L1) + U01 + L0 * Rzr * L1;
        G10 = U10 * d0 + L1 * Rzr * L0;
        G11 = U10 * (-L1) + U11 + L1 * Rzr * L1;
        double gres = gm - g0;
        chi_zr += (gres / (G00 + Rzr)) * gres;
    }
    double dof = (double)(n - 2);
    pval_xy = gtf_gammq(0.5 * dof, 0.5 * chi_xy);
    pval_zr = gtf_gammq(0.5 * dof, 0.5 * chi_zr);
}

// ------------------------------------------------------------------------------------------------
// Information form of a block Gaussian, used by the greedy merge loop of the clustering kernel:
// S = Sigma^-1 (2x2 block + scalar), v = S mu.  Merging is then a plain sum (clustering.py:97-105 computes
// Sigma_m = (S1 + S2)^-1, mu_m = Sigma_m (S1 mu1 + S2 mu2)), and KL-to-merged needs no further inverse because
// inv(Sigma_m) is the running sum itself.  One 2x2 inverse + one reciprocal per round instead of five.
struct GtfInfo {
    double s00, s01, s11, sq, v0, v1, vc, vt;
};
GTF_HD void gtf_to_info(const GtfState &s, GtfInfo &I)
{
    // 1/det and 1/p22 from one reciprocal of their product
    double det = s.p00 * s.p11 - s.p01 * s.p01;
    double r = 1.0 / (det * s.p22);
    double id = r * s.p22;
    I.s00 = s.p11 * id;
    I.s01 = -s.p01 * id;
    I.s11 = s.p00 * id;
    I.sq = r * det;
    I.v0 = I.s00 * s.a + I.s01 * s.b;
    I.v1 = I.s01 * s.a + I.s11 * s.b;
    I.vc = I.sq * s.c;
    I.vt = I.sq * s.tau;
}
GTF_HD void gtf_info_add(GtfInfo &m, const GtfInfo &e)
{
    m.s00 += e.s00; m.s01 += e.s01; m.s11 += e.s11; m.sq += e.sq;
    m.v0 += e.v0; m.v1 += e.v1; m.vc += e.vc; m.vt += e.vt;
}
GTF_HD void gtf_from_info(const GtfInfo &I, GtfState &s)
{
    // 1/det and 1/sq from one reciprocal of their product
    double det = I.s00 * I.s11 - I.s01 * I.s01;
    double r = 1.0 / (det * I.sq);
    double id = r * I.sq;
    s.p00 = I.s11 * id;
    s.p01 = -I.s01 * id;
    s.p11 = I.s00 * id;
    s.p22 = r * det;
    s.a = s.p00 * I.v0 + s.p01 * I.v1;
    s.b = s.p01 * I.v0 + s.p11 * I.v1;
    s.c = s.p22 * I.vc;
    s.tau = s.p22 * I.vt;
}
// KLDistance(entry, merged) with both inverses already at hand (element-wise trace, clustering.py:90-94)
GTF_HD double gtf_kl_info(const GtfState &e, const GtfInfo &ei, const GtfState &m, const GtfInfo &mi)
{
    double tr = (e.p00 - m.p00) * (mi.s00 - ei.s00) + (e.p11 - m.p11) * (mi.s11 - ei.s11) + (e.p22 - m.p22) * (mi.sq - ei.sq);
    double d0 = e.a - m.a, d1 = e.b - m.b, d2 = e.tau - m.tau;
    return tr + (d0 * d0 * (ei.s00 + mi.s00) + 2.0 * d0 * d1 * (ei.s01 + mi.s01) + d1 * d1 * (ei.s11 + mi.s11) +
                 d2 * d2 * (ei.sq + mi.sq));
}
