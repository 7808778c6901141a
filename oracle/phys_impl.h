/*
 * TEST INFRASTRUCTURE — CPU oracle for the physics substep.
 *
 * PARITY UNPINNED: the reference's physics is the closed-source PhysX 5 inside
 * Isaac Gym preview 4 (IsaacGym_Preview_4_Package/.../_bindings/linux-x86_64/
 * libcarb.gym.plugin.so, libPhysXGpu_64.so); it cannot be run, read or compiled
 * here and the reference holds no golden vectors for it.  This file therefore
 * states OUR dynamics spec ("GRX-dyn v1") in plain C; the CUDA
 * kernels (wiki-grx-gym_b200/csrc/grx_env.cu for the lower-limb tree,
 * csrc/grx_phys_generic.cu for any tree incl. self-collision) must match it to fp32 tolerance.
 * The reference call sites this spec stands behind:
 *   legged_robot_fftai.py:51-88  (substep loop, action delay, averages)
 *   legged_robot.py:679-715      (_compute_torques, PD law + motor strength + clip)
 *   legged_robot_fftai.py:67-76  (set_dof_actuation_force_tensor / simulate / refresh_*)
 * Parameters honoured from legged_robot_config.py:35-52: dt, gravity,
 * contact_offset, bounce_threshold_velocity, max_depenetration_velocity,
 * num_position_iterations (used as the PGS sweep count), friction/restitution
 * combine = average (Isaac Gym release notes, docs/release-notes.rst.txt:26).
 *
 * Included twice by phys_oracle.c with REAL = float / double and SFX = _f32/_f64.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may load the library built from this file.
 *
 * Generalised velocity order (internal): [ joint rates (nd) | base linear (3, world,
 * at base-link origin) | base angular (3, world) ].  Equations of motion by
 * Kane's method, mass matrix by composite rigid bodies about the base origin,
 * dense Cholesky, velocity-level projected Gauss-Seidel over contacts (normal +
 * 2 friction rows, pyramid friction) and joint-limit rows, semi-implicit Euler.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

typedef struct {
    int nb, nd, nl, ns, nf;
    const int *parent;                 /* [nb] */
    const REAL *jpos, *jrot, *axis;    /* [nb*3] [nb*9] [nb*3] */
    const REAL *mass, *com, *inertia;  /* [nb] [nb*3] [nb*6] (xx yy zz xy xz yz about COM) */
    const REAL *dof_lower, *dof_upper, *dof_vel_limit, *dof_effort; /* [nd] */
    const int *link_body;              /* [nl] */
    const REAL *link_pos, *link_rot;   /* [nl*3] [nl*9] */
    const int *sph_body, *sph_link;    /* [ns] */
    const REAL *sph_pos, *sph_rad;     /* [ns*3] [ns] */
    const int *foot_links;             /* [nf] URDF-link indices of the feet */
    const REAL *kp, *kd, *default_pos; /* [nd] PD gains + default joint angles */
    int npairs;                        /* robot self-collision (legged_robot_config.py:121 self_collisions = 0 = enabled; create_actor(...,
                                        * collision_filter = 0), legged_robot.py:1022-1028): candidate sphere pairs, in priority order */
    const int *pair_a, *pair_b;        /* [npairs] sphere indices (contact-priority order), bodies differ and are not parent / child */
} FN(Model);

typedef struct {
    int type;                    /* 0 plane z=0, 1 heightfield, 2 structured trimesh (heightfield + snapped vertices) */
    int rows, cols;              /* heightfield samples [rows, cols], x = row axis */
    const short *heights;        /* int16 */
    REAL hscale, vscale, border; /* world x = row*hscale - border */
    REAL friction, restitution;
    const signed char *moves;    /* type 2: [rows, cols, 2] vertex shifts (cells) of the reference's steep-edge snapping, terrain_utils.py:315-328 */
    const unsigned char *near_moved; /* type 2: [rows, cols] != 0 where a vertex within the 3 x 3 cells around cell (i, j) is shifted */
} FN(Terrain);

typedef struct {
    REAL dt;
    REAL gravity;           /* -9.81 along z */
    REAL contact_offset;    /* 0.01 */
    REAL bounce_threshold;  /* 0.5 */
    REAL max_depen_vel;     /* 1.0 */
    REAL erp;               /* 0.2: fraction of penetration corrected per substep */
    int solver_iters;       /* 4 */
    int decimation;         /* 10 */
    REAL action_scale;      /* 1.0 */
    int max_contacts;       /* 8: with <= 7 limit rows the solver has at most 31 rows = one warp lane per row + one lane for the unconstrained update */
    int max_self_contacts;  /* 0 = robot self-collision off (the lower-limb fast kernel), else at most this many sphere-sphere contacts per substep */
} FN(SimCfg);

#define MAXB 36
#define MAXV 40
#define MAXC 16
#define MAXSC 8
#define MAXLIM 7
#define MAXROWS (3 * (MAXC + MAXSC) + MAXV)

static inline void FN(v3cross)(const REAL *a, const REAL *b, REAL *o) {
    REAL x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline REAL FN(v3dot)(const REAL *a, const REAL *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void FN(m3v)(const REAL *R, const REAL *v, REAL *o) {
    REAL x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    REAL y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    REAL z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline void FN(m3m)(const REAL *A, const REAL *B, REAL *C) {
    REAL t[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    for (int i = 0; i < 9; i++) C[i] = t[i];
}
static inline void FN(quat2mat)(const REAL *q, REAL *R) { /* q = xyzw */
    REAL x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}
static inline void FN(mat2quat)(const REAL *R, REAL *q) {
    REAL tr = R[0] + R[4] + R[8];
    if (tr > 0) { REAL s = SQRT(tr + 1) * 2; q[3] = s / 4; q[0] = (R[7] - R[5]) / s; q[1] = (R[2] - R[6]) / s; q[2] = (R[3] - R[1]) / s; }
    else if (R[0] > R[4] && R[0] > R[8]) { REAL s = SQRT(1 + R[0] - R[4] - R[8]) * 2; q[3] = (R[7] - R[5]) / s; q[0] = s / 4; q[1] = (R[1] + R[3]) / s; q[2] = (R[2] + R[6]) / s; }
    else if (R[4] > R[8]) { REAL s = SQRT(1 + R[4] - R[0] - R[8]) * 2; q[3] = (R[2] - R[6]) / s; q[0] = (R[1] + R[3]) / s; q[1] = s / 4; q[2] = (R[5] + R[7]) / s; }
    else { REAL s = SQRT(1 + R[8] - R[0] - R[4]) * 2; q[3] = (R[3] - R[1]) / s; q[0] = (R[2] + R[6]) / s; q[1] = (R[5] + R[7]) / s; q[2] = s / 4; }
}
/* rotation about unit axis a by angle th: Rodrigues */
static inline void FN(axang2mat)(const REAL *a, REAL th, REAL *R) {
    REAL c = COS(th), s = SIN(th), t = 1 - c;
    R[0] = c + a[0] * a[0] * t;        R[1] = a[0] * a[1] * t - a[2] * s; R[2] = a[0] * a[2] * t + a[1] * s;
    R[3] = a[1] * a[0] * t + a[2] * s; R[4] = c + a[1] * a[1] * t;        R[5] = a[1] * a[2] * t - a[0] * s;
    R[6] = a[2] * a[0] * t - a[1] * s; R[7] = a[2] * a[1] * t + a[0] * s; R[8] = c + a[2] * a[2] * t;
}

/* ---- terrain query: height and unit normal of the piecewise-linear surface under (x, y).
 * Heightfield cells are split along the (i,j)-(i+1,j+1) diagonal, the same split the
 * reference's trimesh conversion uses (isaacgym/terrain_utils.py:333-348). */
/* splitmix64 finaliser: item hash of the active-set signature (same function in grx_env.cu) */
static inline unsigned long long FN(mix64)(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline void FN(terrain_query)(const FN(Terrain) *t, REAL x, REAL y, REAL *h, REAL *n, int *cell) {
    if (t->type == 0) { *h = 0; n[0] = 0; n[1] = 0; n[2] = 1; cell[0] = cell[1] = cell[2] = 0; return; }
    REAL gx = (x + t->border) / t->hscale, gy = (y + t->border) / t->hscale;
    REAL fi = FLOOR(gx), fj = FLOOR(gy);
    int i = (int)fi, j = (int)fj;
    if (i < 0) { i = 0; gx = 0; } if (j < 0) { j = 0; gy = 0; }
    if (i > t->rows - 2) { i = t->rows - 2; gx = (REAL)(t->rows - 1); }
    if (j > t->cols - 2) { j = t->cols - 2; gy = (REAL)(t->cols - 1); }
    REAL fx = gx - (REAL)i, fy = gy - (REAL)j;
    REAL h00 = t->heights[i * t->cols + j] * t->vscale, h01 = t->heights[i * t->cols + j + 1] * t->vscale;
    REAL h10 = t->heights[(i + 1) * t->cols + j] * t->vscale, h11 = t->heights[(i + 1) * t->cols + j + 1] * t->vscale;
    REAL dhx, dhy;
    if (fx >= fy) { dhx = h10 - h00; dhy = h11 - h10; }
    else          { dhx = h11 - h01; dhy = h01 - h00; }
    cell[0] = i; cell[1] = j; cell[2] = fx >= fy ? 0 : 1;
    *h = h00 + dhx * fx + dhy * fy;
    REAL sx = -dhx / t->hscale, sy = -dhy / t->hscale;
    REAL inv = 1 / SQRT(sx * sx + sy * sy + 1);
    n[0] = sx * inv; n[1] = sy * inv; n[2] = inv;
    if (t->type != 2 || !t->near_moved[i * t->cols + j]) return;
    /* Structured trimesh (gym.add_triangle_mesh of terrain_utils.convert_heightfield_to_trimesh, terrain_utils.py:286-350): vertices next to a
     * steep edge are shifted sideways by one cell, which turns the ramp of the heightfield into a flat tread + a vertical wall.  The top
     * surface under (x, y): of the triangles of the 3 x 3 cells around the nominal cell whose PROJECTION contains the point (vertical walls
     * project to nothing), the highest one.  Vertex (vi, vj) sits at grid coordinates (vi + mx, vj + my). */
    REAL best = (REAL)-1e30;
    for (int di = -1; di <= 1; di++) for (int dj = -1; dj <= 1; dj++) {
        const int ci = i + di, cj = j + dj;
        if (ci < 0 || cj < 0 || ci > t->rows - 2 || cj > t->cols - 2) continue;
        REAL P[4][3];   /* P00 P10 P01 P11 */
        for (int v = 0; v < 4; v++) {
            const int vi = ci + (v & 1), vj = cj + (v >> 1), id = vi * t->cols + vj;
            P[v][0] = (REAL)(vi + t->moves[2 * id]); P[v][1] = (REAL)(vj + t->moves[2 * id + 1]); P[v][2] = t->heights[id] * t->vscale;
        }
        for (int k = 0; k < 2; k++) {   /* k = 0: (P00, P11, P01), k = 1: (P00, P10, P11)  (terrain_utils.py:342-347) */
            const REAL *a = P[0], *b = k ? P[1] : P[3], *c = k ? P[3] : P[2];
            const REAL e1x = b[0] - a[0], e1y = b[1] - a[1], e2x = c[0] - a[0], e2y = c[1] - a[1];
            const REAL det = e1x * e2y - e2x * e1y;
            if (FABS(det) < (REAL)1e-6) continue;
            const REAL px = gx - a[0], py = gy - a[1];
            const REAL u = (px * e2y - e2x * py) / det, w = (e1x * py - px * e1y) / det;
            if (u < (REAL)-1e-5 || w < (REAL)-1e-5 || u + w > (REAL)1.00001) continue;
            const REAL hh = a[2] + u * (b[2] - a[2]) + w * (c[2] - a[2]);
            if (hh > best) {
                best = hh;
                REAL nx = (e1y * (c[2] - a[2]) - (b[2] - a[2]) * e2y) * t->hscale, ny = ((b[2] - a[2]) * e2x - e1x * (c[2] - a[2])) * t->hscale,
                     nz = det * t->hscale * t->hscale;
                if (nz < 0) { nx = -nx; ny = -ny; nz = -nz; }
                const REAL il = 1 / SQRT(nx * nx + ny * ny + nz * nz);
                *h = hh; n[0] = nx * il; n[1] = ny * il; n[2] = nz * il;
                cell[0] = ci; cell[1] = cj; cell[2] = 2 + k;
            }
        }
    }
}

typedef struct {
    REAL R[MAXB][9], o[MAXB][3], a[MAXB][3], c[MAXB][3], Iw[MAXB][6];
    REAL w[MAXB][3], vo[MAXB][3], al[MAXB][3], ao[MAXB][3];
    REAL m[MAXB];
} FN(Kin);

/* forward kinematics + body velocities + velocity-product ("bias") accelerations */
static void FN(kinematics)(const FN(Model) *M, const REAL *binert, const REAL *root, const REAL *q, const REAL *qd, FN(Kin) *K) {
    for (int b = 0; b < M->nb; b++) {
        const REAL *com, *I6; REAL mass;
        if (b == 0) {
            FN(quat2mat)(root + 3, K->R[0]);
            for (int k = 0; k < 3; k++) { K->o[0][k] = root[k]; K->a[0][k] = 0; K->vo[0][k] = root[7 + k]; K->w[0][k] = root[10 + k]; K->al[0][k] = 0; K->ao[0][k] = 0; }
            mass = binert[0]; com = binert + 1; I6 = binert + 4;
        } else {
            int p = M->parent[b];
            REAL Rj[9], Rq[9], r[3], t1[3], t2[3];
            FN(m3m)(K->R[p], M->jrot + 9 * b, Rj);          /* joint frame in world at q = 0 */
            FN(axang2mat)(M->axis + 3 * b, q[b - 1], Rq);
            FN(m3m)(Rj, Rq, K->R[b]);
            FN(m3v)(K->R[p], M->jpos + 3 * b, r);           /* r = o_b - o_p */
            for (int k = 0; k < 3; k++) K->o[b][k] = K->o[p][k] + r[k];
            FN(m3v)(K->R[b], M->axis + 3 * b, K->a[b]);
            /* velocities */
            FN(v3cross)(K->w[p], r, t1);
            for (int k = 0; k < 3; k++) K->vo[b][k] = K->vo[p][k] + t1[k];
            FN(v3cross)(K->al[p], r, t2);
            REAL t3[3]; FN(v3cross)(K->w[p], t1, t3);
            for (int k = 0; k < 3; k++) K->ao[b][k] = K->ao[p][k] + t2[k] + t3[k];
            FN(v3cross)(K->w[p], K->a[b], t1);
            for (int k = 0; k < 3; k++) { K->w[b][k] = K->w[p][k] + K->a[b][k] * qd[b - 1]; K->al[b][k] = K->al[p][k] + t1[k] * qd[b - 1]; }
            mass = M->mass[b]; com = M->com + 3 * b; I6 = M->inertia + 6 * b;
        }
        K->m[b] = mass;
        REAL rc[3]; FN(m3v)(K->R[b], com, rc);
        for (int k = 0; k < 3; k++) K->c[b][k] = K->o[b][k] + rc[k];
        /* world inertia R I R^T */
        REAL I[9] = {I6[0], I6[3], I6[4], I6[3], I6[1], I6[5], I6[4], I6[5], I6[2]}, T[9], Rt[9];
        FN(m3m)(K->R[b], I, T);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rt[3 * i + j] = K->R[b][3 * j + i];
        FN(m3m)(T, Rt, T);
        K->Iw[b][0] = T[0]; K->Iw[b][1] = T[4]; K->Iw[b][2] = T[8]; K->Iw[b][3] = T[1]; K->Iw[b][4] = T[2]; K->Iw[b][5] = T[5];
    }
}
static inline void FN(sym6v)(const REAL *S, const REAL *v, REAL *o) {
    REAL x = S[0] * v[0] + S[3] * v[1] + S[4] * v[2];
    REAL y = S[3] * v[0] + S[1] * v[1] + S[5] * v[2];
    REAL z = S[4] * v[0] + S[5] * v[1] + S[2] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline int FN(is_ancestor_or_self)(const FN(Model) *M, int anc, int b) {
    while (b >= 0) { if (b == anc) return 1; b = M->parent[b]; }
    return 0;
}

/* Mass matrix Mq[nv*nv] (row-major, internal order) and bias vector h[nv] (Coriolis/centrifugal + gravity). */
static void FN(mass_and_bias)(const FN(Model) *M, const FN(SimCfg) *cfg, const FN(Kin) *K, REAL *Mq, REAL *h) {
    const int nb = M->nb, nd = M->nd, nv = nd + 6;
    const REAL *ref = K->o[0];
    /* per-body spatial inertia about ref (mass, first moment, rotational inertia) and bias wrench about ref */
    REAL cm[MAXB], ch[MAXB][3], cI[MAXB][6], wf[MAXB][3], wn[MAXB][3];
    for (int b = 0; b < nb; b++) {
        REAL r[3] = {K->c[b][0] - ref[0], K->c[b][1] - ref[1], K->c[b][2] - ref[2]};
        REAL m = K->m[b], rr = FN(v3dot)(r, r);
        cm[b] = m; for (int k = 0; k < 3; k++) ch[b][k] = m * r[k];
        cI[b][0] = K->Iw[b][0] + m * (rr - r[0] * r[0]); cI[b][1] = K->Iw[b][1] + m * (rr - r[1] * r[1]); cI[b][2] = K->Iw[b][2] + m * (rr - r[2] * r[2]);
        cI[b][3] = K->Iw[b][3] - m * r[0] * r[1]; cI[b][4] = K->Iw[b][4] - m * r[0] * r[2]; cI[b][5] = K->Iw[b][5] - m * r[1] * r[2];
        /* bias wrench: f = m (a_c - g), n_ref = I al + w x I w + r x f */
        REAL rc[3] = {K->c[b][0] - K->o[b][0], K->c[b][1] - K->o[b][1], K->c[b][2] - K->o[b][2]};
        REAL t1[3], t2[3], ac[3], Iw_[3], Ial[3], g3[3];
        FN(v3cross)(K->al[b], rc, t1); FN(v3cross)(K->w[b], rc, t2); FN(v3cross)(K->w[b], t2, t2);
        for (int k = 0; k < 3; k++) ac[k] = K->ao[b][k] + t1[k] + t2[k];
        ac[2] -= cfg->gravity;
        for (int k = 0; k < 3; k++) wf[b][k] = m * ac[k];
        FN(sym6v)(K->Iw[b], K->w[b], Iw_); FN(sym6v)(K->Iw[b], K->al[b], Ial);
        FN(v3cross)(K->w[b], Iw_, g3); FN(v3cross)(r, wf[b], t1);
        for (int k = 0; k < 3; k++) wn[b][k] = Ial[k] + g3[k] + t1[k];
    }
    /* subtree (composite) sums: children have larger indices than parents (DFS order) */
    for (int b = nb - 1; b >= 1; b--) {
        int p = M->parent[b];
        cm[p] += cm[b];
        for (int k = 0; k < 3; k++) { ch[p][k] += ch[b][k]; wf[p][k] += wf[b][k]; wn[p][k] += wn[b][k]; }
        for (int k = 0; k < 6; k++) cI[p][k] += cI[b][k];
    }
    for (int i = 0; i < nv * nv; i++) Mq[i] = 0;
    /* joint columns: S_j = (lin at ref = a x (ref - o_j), ang = a); F_j = Ic_j S_j = (force, moment about ref) */
    REAL Sl[MAXB][3], Sa[MAXB][3], Ff[MAXB][3], Fn[MAXB][3];
    for (int j = 1; j < nb; j++) {
        REAL d[3] = {ref[0] - K->o[j][0], ref[1] - K->o[j][1], ref[2] - K->o[j][2]};
        FN(v3cross)(K->a[j], d, Sl[j]);
        for (int k = 0; k < 3; k++) Sa[j][k] = K->a[j][k];
        /* force = m v + w x h ; moment = h x v + I w */
        REAL t1[3], t2[3], t3[3];
        FN(v3cross)(Sa[j], ch[j], t1);
        for (int k = 0; k < 3; k++) Ff[j][k] = cm[j] * Sl[j][k] + t1[k];
        FN(v3cross)(ch[j], Sl[j], t2); FN(sym6v)(cI[j], Sa[j], t3);
        for (int k = 0; k < 3; k++) Fn[j][k] = t2[k] + t3[k];
        h[j - 1] = FN(v3dot)(Sl[j], wf[j]) + FN(v3dot)(Sa[j], wn[j]);
    }
    for (int j = 1; j < nb; j++) {
        for (int i = j; i >= 1; i = M->parent[i]) { /* i ancestor-or-self of j */
            REAL v = FN(v3dot)(Sl[i], Ff[j]) + FN(v3dot)(Sa[i], Fn[j]);
            Mq[(i - 1) * nv + (j - 1)] = v; Mq[(j - 1) * nv + (i - 1)] = v;
        }
        for (int k = 0; k < 3; k++) {
            Mq[(nd + k) * nv + (j - 1)] = Ff[j][k]; Mq[(j - 1) * nv + nd + k] = Ff[j][k];
            Mq[(nd + 3 + k) * nv + (j - 1)] = Fn[j][k]; Mq[(j - 1) * nv + nd + 3 + k] = Fn[j][k];
        }
    }
    /* base block: [[m 1, -[h]x],[ [h]x, I ]] */
    {
        REAL m = cm[0], *hh = ch[0], *I = cI[0];
        int L = nd, A = nd + 3;
        for (int k = 0; k < 3; k++) Mq[(L + k) * nv + L + k] = m;
        REAL hx[9] = {0, -hh[2], hh[1], hh[2], 0, -hh[0], -hh[1], hh[0], 0};
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
            Mq[(A + i) * nv + L + j] = hx[3 * i + j];
            Mq[(L + j) * nv + A + i] = hx[3 * i + j];
        }
        Mq[(A + 0) * nv + A + 0] = I[0]; Mq[(A + 1) * nv + A + 1] = I[1]; Mq[(A + 2) * nv + A + 2] = I[2];
        Mq[(A + 0) * nv + A + 1] = Mq[(A + 1) * nv + A + 0] = I[3];
        Mq[(A + 0) * nv + A + 2] = Mq[(A + 2) * nv + A + 0] = I[4];
        Mq[(A + 1) * nv + A + 2] = Mq[(A + 2) * nv + A + 1] = I[5];
        for (int k = 0; k < 3; k++) { h[L + k] = wf[0][k]; h[A + k] = wn[0][k]; }
    }
}

static int FN(cholesky)(REAL *A, int n) { /* in place, lower; returns 0 ok */
    for (int k = 0; k < n; k++) {
        REAL d = A[k * n + k];
        for (int p = 0; p < k; p++) d -= A[k * n + p] * A[k * n + p];
        if (!(d > 0)) return 1;
        d = SQRT(d); A[k * n + k] = d;
        for (int i = k + 1; i < n; i++) {
            REAL s = A[i * n + k];
            for (int p = 0; p < k; p++) s -= A[i * n + p] * A[k * n + p];
            A[i * n + k] = s / d;
        }
    }
    return 0;
}
static void FN(chol_solve)(const REAL *L, int n, REAL *x) {
    for (int i = 0; i < n; i++) { REAL s = x[i]; for (int p = 0; p < i; p++) s -= L[i * n + p] * x[p]; x[i] = s / L[i * n + i]; }
    for (int i = n - 1; i >= 0; i--) { REAL s = x[i]; for (int p = i + 1; p < n; p++) s -= L[p * n + i] * x[p]; x[i] = s / L[i * n + i]; }
}

/* Jacobian row of world point x on body b along direction d (internal velocity order). */
static void FN(point_jac_row)(const FN(Model) *M, const FN(Kin) *K, int b, const REAL *x, const REAL *d, REAL *J) {
    const int nd = M->nd, nv = nd + 6;
    for (int i = 0; i < nv; i++) J[i] = 0;
    for (int j = b; j >= 1; j = M->parent[j]) {
        REAL r[3] = {x[0] - K->o[j][0], x[1] - K->o[j][1], x[2] - K->o[j][2]}, t[3];
        FN(v3cross)(K->a[j], r, t);
        J[j - 1] = FN(v3dot)(t, d);
    }
    REAL r[3] = {x[0] - K->o[0][0], x[1] - K->o[0][1], x[2] - K->o[0][2]}, t[3];
    FN(v3cross)(r, d, t);
    for (int k = 0; k < 3; k++) { J[nd + k] = d[k]; J[nd + 3 + k] = t[k]; }
}

typedef struct {
    int link[MAXC + MAXSC], link2[MAXC + MAXSC];   /* link2 >= 0: self-contact, the reaction goes to that link */
    REAL n[MAXC + MAXSC][3], t1[MAXC + MAXSC][3], t2[MAXC + MAXSC][3];
    REAL lam[MAXC + MAXSC][3];
    int count;
} FN(Contacts);

/* One dt: state (root, q, qd) advanced in place given joint torques tau.  cf_out[nl*3] = net contact force per URDF link.
 * sig (nullable): ACTIVE-SET SIGNATURE of the substep = wrapping sum of mix64(item) over the discrete decisions the step takes:
 *   per accepted contact  item = 1<<56 | sphere s | cell i << 6 | cell j << 18 | triangle (0-3) << 30 | bounce branch << 33 | y-tangent << 34
 *   per joint-limit row   item = 2<<56 | joint j | (upper ? 1 : 0) << 6
 * Two implementations that took the same decisions must agree to rounding; a differing signature explains a differing row
 * (tests/test_env_gpu.py::test_full_step_matches_oracle). */
static int FN(substep)(const FN(Model) *M, const FN(Terrain) *T, const FN(SimCfg) *cfg, const REAL *binert, REAL mu_env, REAL rest_env,
                       REAL *root, REAL *q, REAL *qd, const REAL *tau, const FN(Kin) *K, REAL *cf_out, unsigned long long *sig) {
    unsigned long long sg = 0;
    const int nd = M->nd, nv = nd + 6;
    const REAL dt = cfg->dt;
    REAL Mq[MAXV * MAXV], h[MAXV], u[MAXV], rhs[MAXV];
    FN(mass_and_bias)(M, cfg, K, Mq, h);
    if (FN(cholesky)(Mq, nv)) return 1;
    for (int j = 0; j < nd; j++) { u[j] = qd[j]; rhs[j] = tau[j] - h[j]; }
    for (int k = 0; k < 6; k++) { u[nd + k] = root[7 + k]; rhs[nd + k] = -h[nd + k]; }
    FN(chol_solve)(Mq, nv, rhs);
    for (int i = 0; i < nv; i++) u[i] += dt * rhs[i];

    /* ---- constraint rows */
    static const REAL ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0};
    REAL J[MAXROWS][MAXV], Y[MAXROWS][MAXV], Ad[MAXROWS], bias[MAXROWS];
    FN(Contacts) C; C.count = 0;
    REAL mu = (REAL)0.5 * (mu_env + T->friction), rest = (REAL)0.5 * (rest_env + T->restitution);
    for (int s = 0; s < M->ns && C.count < cfg->max_contacts; s++) {
        int b = M->sph_body[s];
        REAL xs[3], hgt, n[3];
        int cell[3];
        FN(m3v)(K->R[b], M->sph_pos + 3 * s, xs);
        for (int k = 0; k < 3; k++) xs[k] += K->o[b][k];
        FN(terrain_query)(T, xs[0], xs[1], &hgt, n, cell);
        REAL d = (xs[2] - hgt) * n[2] - M->sph_rad[s];
        if (!(d < cfg->contact_offset)) continue;
        int c = C.count++;
        C.link[c] = M->sph_link[s]; C.link2[c] = -1;
        REAL xc[3] = {xs[0] - n[0] * M->sph_rad[s], xs[1] - n[1] * M->sph_rad[s], xs[2] - n[2] * M->sph_rad[s]};
        /* tangent frame: t1 = normalised projection of world x (or y if n ~ x) */
        const int usey = (n[0] > (REAL)0.9 || n[0] < (REAL)-0.9);
        const REAL *e = usey ? ey : ex;
        REAL dn = FN(v3dot)(e, n), t1[3] = {e[0] - dn * n[0], e[1] - dn * n[1], e[2] - dn * n[2]};
        REAL inv = 1 / SQRT(FN(v3dot)(t1, t1)); for (int k = 0; k < 3; k++) t1[k] *= inv;
        REAL t2[3]; FN(v3cross)(n, t1, t2);
        for (int k = 0; k < 3; k++) { C.n[c][k] = n[k]; C.t1[c][k] = t1[k]; C.t2[c][k] = t2[k]; C.lam[c][k] = 0; }
        FN(point_jac_row)(M, K, b, xc, n, J[3 * c]);
        FN(point_jac_row)(M, K, b, xc, t1, J[3 * c + 1]);
        FN(point_jac_row)(M, K, b, xc, t2, J[3 * c + 2]);
        /* normal target velocity */
        REAL vn0 = 0; for (int i = 0; i < nv; i++) vn0 += J[3 * c][i] * (i < nd ? qd[i] : root[7 + i - nd]); /* pre-step approach speed */
        REAL target;
        if (d > 0) target = -d / dt;
        else { target = -d * cfg->erp / dt; if (target > cfg->max_depen_vel) target = cfg->max_depen_vel; }
        const int bounce = (vn0 < -cfg->bounce_threshold && -rest * vn0 > target);
        if (bounce) target = -rest * vn0;
        sg += FN(mix64)((1ull << 56) | (unsigned long long)s | ((unsigned long long)cell[0] << 6) | ((unsigned long long)cell[1] << 18) |
                        ((unsigned long long)cell[2] << 30) | ((unsigned long long)bounce << 33) | ((unsigned long long)usey << 34));
        bias[3 * c] = target; bias[3 * c + 1] = 0; bias[3 * c + 2] = 0;
    }
    /* ---- robot self-collision: sphere-sphere contacts between non-adjacent bodies, candidate pairs in priority order; normal from sphere b to
     * sphere a, contact point in the middle of the gap, rows = J_a - J_b (normal + 2 friction, mu = the robot's own material), same target
     * velocity rule as the ground contacts.  Signature item = 3<<56 | pair index. */
    {
        int nself = 0;
        for (int pi = 0; pi < M->npairs && nself < cfg->max_self_contacts; pi++) {
            const int sa = M->pair_a[pi], sb = M->pair_b[pi], ba = M->sph_body[sa], bb = M->sph_body[sb];
            REAL xa[3], xb[3];
            FN(m3v)(K->R[ba], M->sph_pos + 3 * sa, xa); FN(m3v)(K->R[bb], M->sph_pos + 3 * sb, xb);
            for (int k = 0; k < 3; k++) { xa[k] += K->o[ba][k]; xb[k] += K->o[bb][k]; }
            REAL dv[3] = {xa[0] - xb[0], xa[1] - xb[1], xa[2] - xb[2]};
            const REAL dist = SQRT(FN(v3dot)(dv, dv)), d = dist - M->sph_rad[sa] - M->sph_rad[sb];
            if (!(d < cfg->contact_offset) || !(dist > (REAL)1e-9)) continue;
            nself++;
            int c = C.count++;
            C.link[c] = M->sph_link[sa]; C.link2[c] = M->sph_link[sb];
            REAL n[3] = {dv[0] / dist, dv[1] / dist, dv[2] / dist};
            const REAL mid = M->sph_rad[sb] + (REAL)0.5 * d;
            REAL xc[3] = {xb[0] + n[0] * mid, xb[1] + n[1] * mid, xb[2] + n[2] * mid};
            const int usey = (n[0] > (REAL)0.9 || n[0] < (REAL)-0.9);
            const REAL *e = usey ? ey : ex;
            REAL dn = FN(v3dot)(e, n), t1[3] = {e[0] - dn * n[0], e[1] - dn * n[1], e[2] - dn * n[2]};
            REAL inv = 1 / SQRT(FN(v3dot)(t1, t1)); for (int k = 0; k < 3; k++) t1[k] *= inv;
            REAL t2[3]; FN(v3cross)(n, t1, t2);
            for (int k = 0; k < 3; k++) { C.n[c][k] = n[k]; C.t1[c][k] = t1[k]; C.t2[c][k] = t2[k]; C.lam[c][k] = 0; }
            const REAL *dirs[3] = {n, t1, t2};
            for (int k = 0; k < 3; k++) {
                REAL Jb[MAXV];
                FN(point_jac_row)(M, K, ba, xc, dirs[k], J[3 * c + k]);
                FN(point_jac_row)(M, K, bb, xc, dirs[k], Jb);
                for (int i = 0; i < nv; i++) J[3 * c + k][i] -= Jb[i];
            }
            REAL vn0 = 0; for (int i = 0; i < nv; i++) vn0 += J[3 * c][i] * (i < nd ? qd[i] : root[7 + i - nd]);
            REAL target;
            if (d > 0) target = -d / dt;
            else { target = -d * cfg->erp / dt; if (target > cfg->max_depen_vel) target = cfg->max_depen_vel; }
            const int bounce = (vn0 < -cfg->bounce_threshold && -rest_env * vn0 > target);
            if (bounce) target = -rest_env * vn0;
            sg += FN(mix64)((3ull << 56) | (unsigned long long)pi | ((unsigned long long)bounce << 33) | ((unsigned long long)usey << 34));
            bias[3 * c] = target; bias[3 * c + 1] = 0; bias[3 * c + 2] = 0;
        }
    }
    int nrows = 3 * C.count;
    /* joint limit rows (speculative, predicted with the PRE-step joint rate so that all constraint rows are known
     * before the single M^-1 solve pass): lower: qd >= (lo - q)/dt ; upper: -qd >= (q - hi)/dt.  At most MAXLIM rows. */
    int lim_joint[MAXV]; REAL lim_sign[MAXV], lim_lam[MAXV]; int nlim = 0;
    for (int j = 0; j < nd && nlim < MAXLIM; j++) {
        REAL qn = q[j] + dt * qd[j];
        REAL sgn = 0, tgt = 0;
        if (qn < M->dof_lower[j]) { sgn = 1; tgt = (M->dof_lower[j] - q[j]) / dt; }
        else if (qn > M->dof_upper[j]) { sgn = -1; tgt = (q[j] - M->dof_upper[j]) / dt; }
        if (sgn != 0) {
            int r = nrows + nlim;
            for (int i = 0; i < nv; i++) J[r][i] = 0;
            J[r][j] = sgn; bias[r] = tgt;
            lim_joint[nlim] = j; lim_sign[nlim] = sgn; lim_lam[nlim] = 0; nlim++;
            sg += FN(mix64)((2ull << 56) | (unsigned long long)j | ((unsigned long long)(sgn < 0 ? 1 : 0) << 6));
        }
    }
    int ntot = nrows + nlim;
    for (int r = 0; r < ntot; r++) {
        for (int i = 0; i < nv; i++) Y[r][i] = J[r][i];
        FN(chol_solve)(Mq, nv, Y[r]);
        REAL a = 0; for (int i = 0; i < nv; i++) a += J[r][i] * Y[r][i];
        Ad[r] = a;
    }
    /* ---- projected Gauss-Seidel, velocity space */
    for (int it = 0; it < cfg->solver_iters; it++) {
        for (int c = 0; c < C.count; c++) {
            for (int k = 0; k < 3; k++) {
                int r = 3 * c + k;
                REAL v = 0; for (int i = 0; i < nv; i++) v += J[r][i] * u[i];
                REAL dl = -(v - bias[r]) / Ad[r], ln = C.lam[c][k] + dl;
                if (k == 0) { if (ln < 0) ln = 0; }
                else { REAL lim = (C.link2[c] >= 0 ? mu_env : mu) * C.lam[c][0]; if (ln > lim) ln = lim; if (ln < -lim) ln = -lim; }
                dl = ln - C.lam[c][k]; C.lam[c][k] = ln;
                for (int i = 0; i < nv; i++) u[i] += Y[r][i] * dl;
            }
        }
        for (int l = 0; l < nlim; l++) {
            int r = nrows + l;
            REAL v = 0; for (int i = 0; i < nv; i++) v += J[r][i] * u[i];
            REAL dl = -(v - bias[r]) / Ad[r], ln = lim_lam[l] + dl;
            if (ln < 0) ln = 0;
            dl = ln - lim_lam[l]; lim_lam[l] = ln;
            for (int i = 0; i < nv; i++) u[i] += Y[r][i] * dl;
        }
    }
    (void)lim_joint; (void)lim_sign;
    if (sig) *sig = sg;
    /* ---- contact force report (impulse / dt), per URDF link, world frame, force ON the body */
    for (int i = 0; i < M->nl * 3; i++) cf_out[i] = 0;
    for (int c = 0; c < C.count; c++)
        for (int k = 0; k < 3; k++) {
            const REAL f = (C.n[c][k] * C.lam[c][0] + C.t1[c][k] * C.lam[c][1] + C.t2[c][k] * C.lam[c][2]) / dt;
            cf_out[3 * C.link[c] + k] += f;
            if (C.link2[c] >= 0) cf_out[3 * C.link2[c] + k] -= f;   /* reaction on the other link of a self-contact */
        }
    /* ---- joint velocity limit + integrate */
    for (int j = 0; j < nd; j++) {
        REAL v = u[j], vl = M->dof_vel_limit[j];
        if (v > vl) v = vl; if (v < -vl) v = -vl;
        qd[j] = v; q[j] += dt * v;
    }
    for (int k = 0; k < 6; k++) root[7 + k] = u[nd + k];
    for (int k = 0; k < 3; k++) root[k] += dt * root[7 + k];
    { /* q <- dq(w dt) * q, world-frame angular velocity */
        REAL wx = root[10], wy = root[11], wz = root[12];
        REAL wn = SQRT(wx * wx + wy * wy + wz * wz), th = wn * dt;
        REAL s, c = COS((REAL)0.5 * th);
        if (wn > (REAL)1e-9) s = SIN((REAL)0.5 * th) / wn; else s = (REAL)0.5 * dt;
        REAL dq[4] = {wx * s, wy * s, wz * s, c}, *p = root + 3;
        REAL x = dq[3] * p[0] + dq[0] * p[3] + dq[1] * p[2] - dq[2] * p[1];
        REAL y = dq[3] * p[1] - dq[0] * p[2] + dq[1] * p[3] + dq[2] * p[0];
        REAL z = dq[3] * p[2] + dq[0] * p[1] - dq[1] * p[0] + dq[2] * p[3];
        REAL w = dq[3] * p[3] - dq[0] * p[0] - dq[1] * p[1] - dq[2] * p[2];
        REAL nn = 1 / SQRT(x * x + y * y + z * z + w * w);
        p[0] = x * nn; p[1] = y * nn; p[2] = z * nn; p[3] = w * nn;
    }
    return 0;
}

/* world state (pos3, quat4 xyzw, linvel3 at link origin, angvel3) of URDF link l */
static void FN(link_state)(const FN(Model) *M, const FN(Kin) *K, int l, REAL *out) {
    int b = M->link_body[l];
    REAL r[3], R[9], t[3];
    FN(m3v)(K->R[b], M->link_pos + 3 * l, r);
    FN(m3m)(K->R[b], M->link_rot + 9 * l, R);
    FN(mat2quat)(R, out + 3);
    FN(v3cross)(K->w[b], r, t);
    for (int k = 0; k < 3; k++) { out[k] = K->o[b][k] + r[k]; out[7 + k] = K->vo[b][k] + t[k]; out[10 + k] = K->w[b][k]; }
}

/*
 * One policy step of physics for N envs: the body of during_physics_step
 * (legged_robot_fftai.py:51-88) with the simulate() call replaced by substep().
 *   delay: the per-step scalar drawn at legged_robot_fftai.py:53-54 (substeps with deci < delay use last_actions)
 *   outputs: torques (last substep), link_state [N,nl,13] and contact_force [N,nl,3] after the last substep,
 *            avg_foot_force [N,nf], avg_foot_linvel / avg_foot_angvel [N,nf,3] (means of |.| over substeps, FF:79-88)
 */
static int FN(physics_step_impl)(const FN(Model) *M, const FN(Terrain) *T, const FN(SimCfg) *cfg, int N,
                                REAL *root, REAL *dof_pos, REAL *dof_vel,
                                const REAL *actions, const REAL *last_actions, REAL delay,
                                const REAL *motor_strength, const REAL *base_inertial, const REAL *friction, const REAL *restitution,
                                REAL *torques, REAL *link_state, REAL *contact_force,
                                REAL *avg_foot_force, REAL *avg_foot_linvel, REAL *avg_foot_angvel, unsigned long long *active_sig) {
    const int nd = M->nd, nl = M->nl, nf = M->nf;
    int err = 0;
    if (M->nb > MAXB || nd + 6 > MAXV || cfg->max_contacts > MAXC || cfg->max_self_contacts > MAXSC) return 2;
#pragma omp parallel for schedule(static) reduction(| : err)
    for (int e = 0; e < N; e++) {
        REAL *rt = root + 13 * e, *q = dof_pos + nd * e, *qd = dof_vel + nd * e, *tq = torques + nd * e;
        const REAL *bin = base_inertial + 10 * e;
        REAL cf[MAXB * 4 * 3]; /* nl <= 48 */
        REAL ls[13];
        FN(Kin) K;
        for (int f = 0; f < nf; f++) { avg_foot_force[e * nf + f] = 0; for (int k = 0; k < 3; k++) { avg_foot_linvel[(e * nf + f) * 3 + k] = 0; avg_foot_angvel[(e * nf + f) * 3 + k] = 0; } }
        FN(kinematics)(M, bin, rt, q, qd, &K);
        for (int deci = 0; deci < cfg->decimation; deci++) {
            const REAL *act = ((REAL)deci < delay ? last_actions : actions) + nd * e;
            for (int j = 0; j < nd; j++) { /* legged_robot.py:691-713 (+ FF:64) */
                REAL t = M->kp[j] * (act[j] * cfg->action_scale + M->default_pos[j] - q[j]) - M->kd[j] * qd[j];
                t *= motor_strength[e * nd + j];
                REAL lim = M->dof_effort[j];
                if (t > lim) t = lim; if (t < -lim) t = -lim;
                tq[j] = t;
            }
            err |= FN(substep)(M, T, cfg, bin, friction[e], restitution[e], rt, q, qd, tq, &K, cf,
                               active_sig ? active_sig + (size_t)e * cfg->decimation + deci : (unsigned long long *)0);
            FN(kinematics)(M, bin, rt, q, qd, &K);
            for (int f = 0; f < nf; f++) {
                int l = M->foot_links[f];
                FN(link_state)(M, &K, l, ls);
                avg_foot_force[e * nf + f] += SQRT(cf[3 * l] * cf[3 * l] + cf[3 * l + 1] * cf[3 * l + 1] + cf[3 * l + 2] * cf[3 * l + 2]);
                for (int k = 0; k < 3; k++) { avg_foot_linvel[(e * nf + f) * 3 + k] += FABS(ls[7 + k]); avg_foot_angvel[(e * nf + f) * 3 + k] += FABS(ls[10 + k]); }
            }
        }
        for (int f = 0; f < nf; f++) {
            avg_foot_force[e * nf + f] /= (REAL)cfg->decimation;
            for (int k = 0; k < 3; k++) { avg_foot_linvel[(e * nf + f) * 3 + k] /= (REAL)cfg->decimation; avg_foot_angvel[(e * nf + f) * 3 + k] /= (REAL)cfg->decimation; }
        }
        for (int l = 0; l < nl; l++) {
            FN(link_state)(M, &K, l, link_state + (size_t)(e * nl + l) * 13);
            for (int k = 0; k < 3; k++) contact_force[(size_t)(e * nl + l) * 3 + k] = cf[3 * l + k];
        }
    }
    return err;
}

int FN(grx_oracle_physics_step)(const FN(Model) *M, const FN(Terrain) *T, const FN(SimCfg) *cfg, int N,
                                REAL *root, REAL *dof_pos, REAL *dof_vel,
                                const REAL *actions, const REAL *last_actions, REAL delay,
                                const REAL *motor_strength, const REAL *base_inertial, const REAL *friction, const REAL *restitution,
                                REAL *torques, REAL *link_state, REAL *contact_force,
                                REAL *avg_foot_force, REAL *avg_foot_linvel, REAL *avg_foot_angvel) {
    return FN(physics_step_impl)(M, T, cfg, N, root, dof_pos, dof_vel, actions, last_actions, delay, motor_strength, base_inertial, friction,
                                 restitution, torques, link_state, contact_force, avg_foot_force, avg_foot_linvel, avg_foot_angvel,
                                 (unsigned long long *)0);
}
/* same, also reporting the active-set signature of every substep: active_sig [N, decimation] */
int FN(grx_oracle_physics_step_sig)(const FN(Model) *M, const FN(Terrain) *T, const FN(SimCfg) *cfg, int N,
                                    REAL *root, REAL *dof_pos, REAL *dof_vel,
                                    const REAL *actions, const REAL *last_actions, REAL delay,
                                    const REAL *motor_strength, const REAL *base_inertial, const REAL *friction, const REAL *restitution,
                                    REAL *torques, REAL *link_state, REAL *contact_force,
                                    REAL *avg_foot_force, REAL *avg_foot_linvel, REAL *avg_foot_angvel, unsigned long long *active_sig) {
    return FN(physics_step_impl)(M, T, cfg, N, root, dof_pos, dof_vel, actions, last_actions, delay, motor_strength, base_inertial, friction,
                                 restitution, torques, link_state, contact_force, avg_foot_force, avg_foot_linvel, avg_foot_angvel, active_sig);
}

/* One simulate() call (dt) with given joint torques for N envs: what FakeGym.simulate() of the reference
 * harness calls (oracle/ref_harness), i.e. the stand-in for gym.simulate + refresh_* (legged_robot_fftai.py:67-76). */
int FN(grx_oracle_substep)(const FN(Model) *M, const FN(Terrain) *T, const FN(SimCfg) *cfg, int N,
                           REAL *root, REAL *dof_pos, REAL *dof_vel, const REAL *tau,
                           const REAL *base_inertial, const REAL *friction, const REAL *restitution,
                           REAL *link_state, REAL *contact_force) {
    const int nd = M->nd, nl = M->nl;
    int err = 0;
    if (M->nb > MAXB || nd + 6 > MAXV || cfg->max_contacts > MAXC) return 2;
#pragma omp parallel for schedule(static) reduction(| : err)
    for (int e = 0; e < N; e++) {
        REAL *rt = root + 13 * e, *q = dof_pos + nd * e, *qd = dof_vel + nd * e;
        const REAL *bin = base_inertial + 10 * e;
        REAL cf[MAXB * 4 * 3];
        FN(Kin) K;
        FN(kinematics)(M, bin, rt, q, qd, &K);
        err |= FN(substep)(M, T, cfg, bin, friction[e], restitution[e], rt, q, qd, tau + nd * e, &K, cf, (unsigned long long *)0);
        FN(kinematics)(M, bin, rt, q, qd, &K);
        for (int l = 0; l < nl; l++) {
            FN(link_state)(M, &K, l, link_state + (size_t)(e * nl + l) * 13);
            for (int k = 0; k < 3; k++) contact_force[(size_t)(e * nl + l) * 3 + k] = cf[3 * l + k];
        }
    }
    return err;
}

/* link states only (FK of the current state), for initialisation of rigid_body_states */
int FN(grx_oracle_link_states)(const FN(Model) *M, int N, const REAL *root, const REAL *dof_pos, const REAL *dof_vel,
                               const REAL *base_inertial, REAL *link_state) {
    for (int e = 0; e < N; e++) {
        FN(Kin) K;
        FN(kinematics)(M, base_inertial + 10 * e, root + 13 * e, dof_pos + M->nd * e, dof_vel + M->nd * e, &K);
        for (int l = 0; l < M->nl; l++) FN(link_state)(M, &K, l, link_state + (size_t)(e * M->nl + l) * 13);
    }
    return 0;
}

/* Diagnostics for the invariants tests: mass matrix, bias vector and total energy of one env. */
int FN(grx_oracle_dynamics_terms)(const FN(Model) *M, const FN(SimCfg) *cfg, const REAL *base_inertial,
                                  const REAL *root, const REAL *q, const REAL *qd, REAL *Mq_out, REAL *h_out, REAL *energy_out) {
    FN(Kin) K;
    FN(kinematics)(M, base_inertial, root, q, qd, &K);
    FN(mass_and_bias)(M, cfg, &K, Mq_out, h_out);
    REAL ke = 0, pe = 0;
    for (int b = 0; b < M->nb; b++) {
        REAL rc[3] = {K.c[b][0] - K.o[b][0], K.c[b][1] - K.o[b][1], K.c[b][2] - K.o[b][2]}, t[3], vc[3], Iw_[3];
        FN(v3cross)(K.w[b], rc, t);
        for (int k = 0; k < 3; k++) vc[k] = K.vo[b][k] + t[k];
        FN(sym6v)(K.Iw[b], K.w[b], Iw_);
        ke += (REAL)0.5 * (K.m[b] * FN(v3dot)(vc, vc) + FN(v3dot)(K.w[b], Iw_));
        pe += -K.m[b] * cfg->gravity * K.c[b][2];
    }
    energy_out[0] = ke; energy_out[1] = pe;
    return 0;
}

#undef MAXB
#undef MAXV
#undef MAXC
#undef MAXSC
#undef MAXLIM
#undef MAXROWS
