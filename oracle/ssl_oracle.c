/*
 * ssl_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see ssl_oracle.h).
 *
 * PARITY UNPINNED (no Bullet, no reference golden vectors).
 * Plain C restatement of the step the reference performs through
 *   world_step -> world_step_delta -> btDiscreteDynamicsWorld::stepSimulation
 *   (reference src/sslsim.cpp:307-326)
 * for the scene it builds (reference src/sslsim.cpp:16-28, 44-49, 63-100,
 * 136-203, 238-252), following docs/STEP_SPEC.md.  Bullet semantics are those
 * of Bullet 2.82 (btSequentialImpulseConstraintSolver::setupContactConstraint,
 * solveGroupCacheFriendlyIterations; btDiscreteDynamicsWorld::
 * internalSingleStepSimulation), restated from their published algorithm.
 *
 * Build: gcc -O2 -ffp-contract=off (no fast-math).  All arithmetic is IEEE
 * binary32 with the association order written here.  The compiler contracts
 * nothing; where STEP_SPEC section 0 says fused multiply-add the code calls
 * fmaf() (through mad / nmad / dot2 / dot3 below), and the CUDA kernel, compiled
 * with --fmad=false, calls __fmaf_rn at the same places, so the two agree
 * bit-for-bit ("FMA in both via fmaf").
 */
#include "ssl_oracle.h"
#include <math.h>
#include <pthread.h>
#include <string.h>

/* ---- scene constants (reference src/sslsim.cpp) ------------------------- */
#define BALL_R 0.0215f        /* :17 BALL_DIAM/2 */
#define BALL_M 0.046f         /* :18 */
#define ROBOT_R 0.09f         /* :19 ROBOT_DIAM/2 */
#define ROBOT_M 2.0f          /* :21 */
#define DROP_Z 0.5f           /* :22 */
#define WHEEL_R 0.025f        /* :23 WHEEL_DIAM/2: robot origin rests at this height */
#define ROBOT_ZLO (-0.02f)    /* :89-90 chassis centre +0.0525, half height 0.0725 */
#define ROBOT_ZHI 0.125f
#define WHEEL_RC 0.085f       /* :115 wheel contact radius ROBOT_DIAM/2 - WHEEL_WIDTH/2 */
#define GRAV 9.80665f         /* :240 */
#define CEIL_Z 10.0f          /* :187 */
#define E_BALL_PLANE 0.5f     /* :47 x :141 (product rule) */
#define E_ROBOT_PLANE 0.2f    /* :87 x :141 */
#define E_BALL_ROBOT 0.1f     /* :47 x :87 */
#define E_ROBOT_ROBOT 0.04f   /* :87 x :87 */
#define E_GOAL 0.0f           /* goal bodies keep Bullet's default restitution 0 */
#define MU 0.25f              /* Bullet default friction 0.5 x 0.5 */
/* Bullet 2.82 defaults */
#define MARGIN 0.02f          /* gContactBreakingThreshold */
#define ERP 0.2f
#define ERP2 0.8f
#define SPLIT_THRESH (-0.04f)
#define SLEEP_LIN2 0.64f      /* btRigidBody default linear sleeping threshold 0.8, squared */
#define SLEEP_ANG2 1.0f       /* default angular sleeping threshold 1.0, squared */
#define SLEEP_TIME 2.0f       /* gDeactivationTime */
/* reach of a robot's collision shapes from its axis: chassis radius + the wheel cylinders that stick out of it
 * (reference :81-98: wheel centres at ROBOT_DIAM/2 + WHEEL_WIDTH/2, half width WHEEL_WIDTH/2) */
#define ROBOT_REACH 0.1f
#define FLT_EPS 1.1920929e-07f
#define L2_MAX 1e37f /* squared lateral speed above which the direction falls back like a zero one (overflow guard) */
#define TWO_PI 6.2831855f
/* wheel mounting angles: reference :27-28 rotates by WHEELS_ANGLE about (0,0,-1) */
static const float WSIN[4] = {-0.8660254f, -0.70710677f, 0.70710677f, 0.8660254f};
static const float WCOS[4] = {0.5f, -0.70710677f, -0.70710677f, 0.5f};

#define MAX_SERIAL_ROWS 64
#define MAX_BODY_STATICS 4

/* STEP_SPEC section 0: the four fused forms; everything else is one rounding per operator */
static inline float mad(float a, float b, float c) { return fmaf(a, b, c); }   /* a*b + c */
static inline float nmad(float a, float b, float c) { return fmaf(-a, b, c); } /* c - a*b */
static inline float dot2(float ax, float ay, float bx, float by) { return fmaf(ay, by, ax * bx); }
static inline float dot3(float ax, float ay, float az, float bx, float by, float bz) {
  return fmaf(az, bz, fmaf(ay, by, ax * bx));
}

void orc_params_default(orc_params *p) {
  memset(p, 0, sizeof *p);
  p->flags = 0;
  p->solver_iterations = 10;
  p->kick_reach = 0.015f;
  p->kick_halfwidth = 0.04f;
  p->kick_max_height = 0.03f;
  p->wheel_kp = 24.0f;
  p->wheel_fmax = 4.0f;
  p->robot_yaw_inertia = 0.0081f; /* 1/2 m r^2 */
  p->dribble_kd = 60.0f;
  p->dribble_cd = 8.0f;
  p->dribble_amax = 12.0f;
  p->dribble_reach = 0.03f;
  p->mu_roll = 0.03f;
  p->restitution_threshold = 0.0f; /* Bullet 2.82: none (0.2 is btContactSolverInfo's default from 2.87 on) */
  p->warmstart_factor = 0.85f;     /* btContactSolverInfo::m_warmstartingFactor with SOLVER_USE_WARMSTARTING (Bullet default) */
}

/* ---- counter RNG --------------------------------------------------------- */
static uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
float orc_u01(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
  uint64_t h = mix64(seed);
  h = mix64(h ^ a);
  h = mix64(h ^ b);
  h = mix64(h ^ c);
  return (float)(uint32_t)(h >> 40) * 5.9604645e-08f; /* 2^-24 */
}

/* ---- sin/cos (STEP_SPEC 3.1): Cody-Waite by pi/2, cephes polynomials ------ */
void orc_sincos(float x, float *s, float *c) {
  float kf = rintf(x * 0.63661975f);
  int k = (int)kf;
  float r = nmad(kf, 1.5703125f, x);
  r = nmad(kf, 4.837512969970703125e-4f, r);
  r = nmad(kf, 7.54978995489188216e-8f, r);
  float z = r * r;
  float sp = mad(-1.9515295891e-4f, z, 8.3321608736e-3f);
  sp = mad(sp, z, -1.6666654611e-1f);
  sp = mad(sp * z, r, r);
  float cp = mad(2.443315711809948e-5f, z, -1.388731625493765e-3f);
  cp = mad(cp, z, 4.166664568298827e-2f);
  cp = mad(cp * z, z, mad(-0.5f, z, 1.0f));
  switch (k & 3) {
  case 0: *s = sp; *c = cp; break;
  case 1: *s = cp; *c = -sp; break;
  case 2: *s = -sp; *c = -cp; break;
  default: *s = -cp; *c = sp; break;
  }
}

/* ---- derived field geometry (reference src/sslsim.cpp:145-203, field.h:36-42) */
typedef struct {
  float half_len, half_wid, limit_x, limit_y;
  float goal_half_w, goal_back, goal_height;
  float lo[6][3], hi[6][3];
} field_derived;

static void derive_field(const orc_field *f, field_derived *d) {
  d->half_len = f->field_length / 2;
  d->half_wid = f->field_width / 2;
  d->limit_x = f->field_length / 2 + f->boundary_width + f->referee_width;
  d->limit_y = f->field_width / 2 + f->boundary_width + f->referee_width;
  d->goal_half_w = f->goal_width / 2;
  d->goal_back = d->half_len + f->goal_depth;
  d->goal_height = f->goal_height;
  float ww = f->goal_wall_width, gd = f->goal_depth, gw = f->goal_width, gh = f->goal_height;
  /* -x goal at (-L/2,0,0), identity (reference :198-199) */
  float sx = -d->half_len + (-gd / 2 - ww / 2), shx = gd / 2 + ww / 2;
  float sy = gw / 2 + ww / 2, shy = ww / 2;
  float rx = -d->half_len + (-gd - ww / 2), rhx = ww / 2, rhy = gw / 2;
  float cz = gh / 2, hz = gh / 2;
  /* 0: side wall at -y, 1: side wall at +y, 2: rear wall */
  d->lo[0][0] = sx - shx; d->hi[0][0] = sx + shx; d->lo[0][1] = -sy - shy; d->hi[0][1] = -sy + shy;
  d->lo[1][0] = sx - shx; d->hi[1][0] = sx + shx; d->lo[1][1] = sy - shy;  d->hi[1][1] = sy + shy;
  d->lo[2][0] = rx - rhx; d->hi[2][0] = rx + rhx; d->lo[2][1] = -rhy;      d->hi[2][1] = rhy;
  for (int k = 0; k < 3; k++) { d->lo[k][2] = cz - hz; d->hi[k][2] = cz + hz; }
  /* +x goal: same compound rotated 180 deg about z at (+L/2,0,0) (reference :200-201) */
  for (int k = 0; k < 3; k++) {
    d->lo[3 + k][0] = -d->hi[k][0]; d->hi[3 + k][0] = -d->lo[k][0];
    d->lo[3 + k][1] = -d->hi[k][1]; d->hi[3 + k][1] = -d->lo[k][1];
    d->lo[3 + k][2] = d->lo[k][2];  d->hi[3 + k][2] = d->hi[k][2];
  }
}

/* ---- contact rows -------------------------------------------------------- */
typedef struct {
  float nx, ny, nz, target;
  float tx, ty, tz, push_target;
  int a, b;   /* body indices; b = -1 for a static partner; normal points b -> a */
  float mu;   /* 0 disables the friction row */
  int spin;   /* ball-ground row in spin mode: friction row carries the angular term */
  float acc_n, acc_f, acc_p;
  float dist, l2; /* separation and squared lateral relative velocity at set-up (4.5a: does the contact point persist?) */
  int sid;    /* static partner: 0 ceiling, 1..4 walls, 5..10 goal boxes, 11 ground; -1 for a pair row (key of the warm table) */
} row_t;

static float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* btPlaneSpace1: first tangent of the plane orthogonal to n */
static void plane_space1(float nx, float ny, float nz, float *tx, float *ty, float *tz) {
  if (fabsf(nz) > 0.70710677f) {
    float a = dot2(ny, nz, ny, nz);
    float k = 1.0f / sqrtf(a);
    *tx = 0.0f; *ty = -nz * k; *tz = ny * k;
  } else {
    float a = dot2(nx, ny, nx, ny);
    float k = 1.0f / sqrtf(a);
    *tx = -ny * k; *ty = nx * k; *tz = 0.0f;
  }
}

typedef struct {
  int n, R;
  float inv_h, h;
  const orc_vec4 *pos, *vel; /* substep-start state (vel after kick) */
  const float *hvx, *hvy, *hvz; /* vel + external impulses */
  int spin_mode;
  float ball_margin;
  float rest_thresh;
  const uint32_t *cflags;    /* per robot command flags or NULL */
} gen_ctx;

/* finish a row given (n, dist, a, b, e, mu): Bullet 2.82 setupContactConstraint */
static void finish_row(const gen_ctx *g, row_t *r, float nx, float ny, float nz, float dist,
                       int a, int b, float e, float mu) {
  r->nx = nx; r->ny = ny; r->nz = nz; r->a = a; r->b = b; r->mu = mu; r->spin = 0; r->sid = -1; r->dist = dist;
  r->acc_n = r->acc_f = r->acc_p = 0.0f;
  float rvx = g->vel[a].x, rvy = g->vel[a].y, rvz = g->vel[a].z;
  if (b >= 0) { rvx = rvx - g->vel[b].x; rvy = rvy - g->vel[b].y; rvz = rvz - g->vel[b].z; }
  float vn = dot3(nx, ny, nz, rvx, rvy, rvz);
  float rest = (-vn > g->rest_thresh) ? e * (-vn) : 0.0f;
  float push = 0.0f;
  if (dist > 0.0f) {
    /* speculative row: the gap may close but not overshoot; if it closes within
     * this substep and the pair is restitutive, reflect now (STEP_SPEC 4.3) */
    float hx = g->hvx[a], hy = g->hvy[a], hz = g->hvz[a];
    if (b >= 0) { hx = hx - g->hvx[b]; hy = hy - g->hvy[b]; hz = hz - g->hvz[b]; }
    float hn = dot3(nx, ny, nz, hx, hy, hz);
    int closing = mad(hn, g->h, dist) < 0.0f;
    r->target = (closing && rest > 0.0f) ? rest : -(dist * g->inv_h);
  } else if (dist > SPLIT_THRESH) {
    r->target = mad(-(dist * ERP), g->inv_h, rest);
  } else {
    r->target = rest;
    push = -(dist * ERP2) * g->inv_h;
  }
  r->push_target = push;
  /* single friction direction along the lateral relative velocity */
  float lx = nmad(nx, vn, rvx), ly = nmad(ny, vn, rvy), lz = nmad(nz, vn, rvz);
  float l2 = dot3(lx, ly, lz, lx, ly, lz);
  r->l2 = l2;
  if (l2 > FLT_EPS && l2 < L2_MAX) {
    float k = 1.0f / sqrtf(l2);
    r->tx = lx * k; r->ty = ly * k; r->tz = lz * k;
  } else {
    plane_space1(nx, ny, nz, &r->tx, &r->ty, &r->tz);
  }
}

static int is_ball(const gen_ctx *g, int b) { return b == g->R; }

/* ground row of body b (always evaluated; one row stands for the robot's four
 * wheel contacts: identical linear Jacobians, zero inverse inertia about x,y) */
static int gen_ground(const gen_ctx *g, int b, row_t *r) {
  int ball = is_ball(g, b);
  float dist = g->pos[b].z - (ball ? BALL_R : WHEEL_R);
  float margin = ball ? g->ball_margin : MARGIN;
  if (!(dist <= margin)) return 0;
  float mu = MU;
  if (!ball && g->cflags && (g->cflags[b] & ORC_CMD_DRIVE)) mu = 0.0f; /* wheels replace sliding friction */
  finish_row(g, r, 0.0f, 0.0f, 1.0f, dist, b, -1, ball ? E_BALL_PLANE : E_ROBOT_PLANE, mu);
  r->sid = 11;
  if (ball && g->spin_mode) {
    /* friction direction from the contact-point velocity (v + w x r_c), r_c = (0,0,-r) */
    float cvx = nmad(BALL_R, g->vel[b].w, g->vel[b].x); /* vel.w = omega_y */
    float cvy = mad(BALL_R, g->pos[b].w, g->vel[b].y);  /* pos.w = omega_x */
    float l2 = dot2(cvx, cvy, cvx, cvy);
    r->l2 = l2; /* the contact point of a rolling ball does not slide */
    if (l2 > FLT_EPS && l2 < L2_MAX) {
      float k = 1.0f / sqrtf(l2);
      r->tx = cvx * k; r->ty = cvy * k; r->tz = 0.0f;
    } else {
      plane_space1(0.0f, 0.0f, 1.0f, &r->tx, &r->ty, &r->tz);
    }
    r->spin = 1;
  }
  return 1;
}

/* static rows other than the ground for body b, canonical order:
 * ceiling, wall -x, wall +x, wall -y, wall +y, goal boxes 0..5 */
static int gen_statics(const gen_ctx *g, const field_derived *d, int b, row_t *out, int cap) {
  int cnt = 0;
  int ball = is_ball(g, b);
  float px = g->pos[b].x, py = g->pos[b].y, pz = g->pos[b].z;
  float rad = ball ? BALL_R : ROBOT_R;
  float margin = ball ? g->ball_margin : MARGIN;
  float e = ball ? E_BALL_PLANE : E_ROBOT_PLANE;
  float dist;
  int sid = 0; /* canonical index of the static solid being tested */
#define EMIT(NX, NY, NZ, DIST, E)                                                     \
  do {                                                                                \
    if (cnt < cap) { finish_row(g, &out[cnt], (NX), (NY), (NZ), (DIST), b, -1, (E), MU); out[cnt].sid = sid; } \
    cnt++;                                                                            \
  } while (0)
  dist = CEIL_Z - (pz + (ball ? BALL_R : ROBOT_ZHI));
  if (dist <= margin) EMIT(0.0f, 0.0f, -1.0f, dist, e);
  dist = (px + d->limit_x) - rad; sid = 1;
  if (dist <= margin) EMIT(1.0f, 0.0f, 0.0f, dist, e);
  dist = (d->limit_x - px) - rad; sid = 2;
  if (dist <= margin) EMIT(-1.0f, 0.0f, 0.0f, dist, e);
  dist = (py + d->limit_y) - rad; sid = 3;
  if (dist <= margin) EMIT(0.0f, 1.0f, 0.0f, dist, e);
  dist = (d->limit_y - py) - rad; sid = 4;
  if (dist <= margin) EMIT(0.0f, -1.0f, 0.0f, dist, e);
  for (int k = 0; k < 6; k++) {
    sid = 5 + k;
    const float *lo = d->lo[k], *hi = d->hi[k];
    float nx, ny, nz;
    if (ball) {
      float cx = clampf(px, lo[0], hi[0]), cy = clampf(py, lo[1], hi[1]), cz = clampf(pz, lo[2], hi[2]);
      float dx = px - cx, dy = py - cy, dz = pz - cz;
      float d2 = dot3(dx, dy, dz, dx, dy, dz);
      if (d2 > 0.0f) {
        float dd = sqrtf(d2);
        nx = dx / dd; ny = dy / dd; nz = dz / dd;
        dist = dd - BALL_R;
      } else { /* centre inside the box: minimum-penetration face, order -x +x -y +y -z +z */
        float pen = px - lo[0]; nx = -1.0f; ny = 0.0f; nz = 0.0f;
        float q = hi[0] - px; if (q < pen) { pen = q; nx = 1.0f; ny = 0.0f; nz = 0.0f; }
        q = py - lo[1]; if (q < pen) { pen = q; nx = 0.0f; ny = -1.0f; nz = 0.0f; }
        q = hi[1] - py; if (q < pen) { pen = q; nx = 0.0f; ny = 1.0f; nz = 0.0f; }
        q = pz - lo[2]; if (q < pen) { pen = q; nx = 0.0f; ny = 0.0f; nz = -1.0f; }
        q = hi[2] - pz; if (q < pen) { pen = q; nx = 0.0f; ny = 0.0f; nz = 1.0f; }
        dist = -pen - BALL_R;
      }
    } else {
      float cx = clampf(px, lo[0], hi[0]), cy = clampf(py, lo[1], hi[1]);
      float dx = px - cx, dy = py - cy;
      float d2 = dot2(dx, dy, dx, dy);
      float hx, hy, sh;
      if (d2 > 0.0f) {
        float dd = sqrtf(d2);
        hx = dx / dd; hy = dy / dd;
        sh = dd - ROBOT_R;
      } else {
        float pen = px - lo[0]; hx = -1.0f; hy = 0.0f;
        float q = hi[0] - px; if (q < pen) { pen = q; hx = 1.0f; hy = 0.0f; }
        q = py - lo[1]; if (q < pen) { pen = q; hx = 0.0f; hy = -1.0f; }
        q = hi[1] - py; if (q < pen) { pen = q; hx = 0.0f; hy = 1.0f; }
        sh = -pen - ROBOT_R;
      }
      float s_up = (pz + ROBOT_ZLO) - hi[2]; /* robot above the box */
      float s_dn = lo[2] - (pz + ROBOT_ZHI); /* robot below the box */
      float sv = s_up >= s_dn ? s_up : s_dn;
      if (sh >= sv) { nx = hx; ny = hy; nz = 0.0f; dist = sh; }
      else { nx = 0.0f; ny = 0.0f; nz = s_up >= s_dn ? 1.0f : -1.0f; dist = sv; }
    }
    if (dist <= margin) EMIT(nx, ny, nz, dist, E_GOAL);
  }
#undef EMIT
  return cnt;
}

/* robot i vs robot j (i < j): parallel capped cylinders; normal points j -> i */
static int gen_robot_robot(const gen_ctx *g, int i, int j, row_t *r) {
  float dx = g->pos[i].x - g->pos[j].x, dy = g->pos[i].y - g->pos[j].y;
  float d2 = dot2(dx, dy, dx, dy);
  float lim = 2.0f * ROBOT_R + MARGIN;
  if (!(d2 <= lim * lim)) return 0;
  float dd = sqrtf(d2);
  float hx = 1.0f, hy = 0.0f;
  if (d2 > 1e-12f) { hx = dx / dd; hy = dy / dd; }
  float sh = dd - 2.0f * ROBOT_R;
  float zi = g->pos[i].z, zj = g->pos[j].z;
  float s_up = (zi + ROBOT_ZLO) - (zj + ROBOT_ZHI); /* i above j */
  float s_dn = (zj + ROBOT_ZLO) - (zi + ROBOT_ZHI);
  float sv = s_up >= s_dn ? s_up : s_dn;
  float nx, ny, nz, dist;
  if (sh >= sv) { nx = hx; ny = hy; nz = 0.0f; dist = sh; }
  else { nx = 0.0f; ny = 0.0f; nz = s_up >= s_dn ? 1.0f : -1.0f; dist = sv; }
  if (!(dist <= MARGIN)) return 0;
  finish_row(g, r, nx, ny, nz, dist, i, j, E_ROBOT_ROBOT, MU);
  return 1;
}

/* ball vs robot i: sphere vs capped cylinder, exact closest feature; normal robot -> ball */
static int gen_ball_robot(const gen_ctx *g, int i, row_t *r) {
  int B = g->R;
  float dx = g->pos[B].x - g->pos[i].x, dy = g->pos[B].y - g->pos[i].y;
  float d2 = dot2(dx, dy, dx, dy);
  float lim = ROBOT_R + BALL_R + g->ball_margin;
  if (!(d2 <= lim * lim)) return 0;
  float dd = sqrtf(d2);
  float hx = 1.0f, hy = 0.0f;
  if (d2 > 1e-12f) { hx = dx / dd; hy = dy / dd; }
  float sh = dd - ROBOT_R;
  float qz = g->pos[B].z - g->pos[i].z;
  float s_lo = ROBOT_ZLO - qz, s_hi = qz - ROBOT_ZHI;
  float sv = s_hi >= s_lo ? s_hi : s_lo;
  float sg = s_hi >= s_lo ? 1.0f : -1.0f;
  float nx, ny, nz, dist;
  if (sh > 0.0f && sv > 0.0f) {
    float ee = sqrtf(dot2(sh, sv, sh, sv));
    float kh = sh / ee, kv = sv / ee;
    nx = hx * kh; ny = hy * kh; nz = sg * kv;
    dist = ee - BALL_R;
  } else if (sh >= sv) { nx = hx; ny = hy; nz = 0.0f; dist = sh - BALL_R; }
  else { nx = 0.0f; ny = 0.0f; nz = sg; dist = sv - BALL_R; }
  if (!(dist <= g->ball_margin)) return 0;
  finish_row(g, r, nx, ny, nz, dist, B, i, E_BALL_ROBOT, MU);
  return 1;
}

/* ---- ball activation state (Bullet btCollisionObject activation states; robots are DISABLE_DEACTIVATION,
 * reference :113, the ball is not, :51-61).  Kept in the status word: bits 18-19 the state, bits 20-31 the
 * number of consecutive slow substeps (Bullet's m_deactivationTime is that many additions of h in binary32, so
 * the count determines it exactly; orc_sleep_substeps gives the count at which it first exceeds 2 s). ------ */
#define ACT_SHIFT 18
#define ACT_ACTIVE 0u  /* ACTIVE_TAG */
#define ACT_WANTS 1u   /* WANTS_DEACTIVATION */
#define ACT_ASLEEP 2u  /* ISLAND_SLEEPING */
#define CNT_SHIFT 20
#define CNT_MAX 4095u

int orc_sleep_substeps(float h) {
  float t = 0.0f;
  for (int n = 1; n <= (int)CNT_MAX; n++) {
    t = t + h;                       /* m_deactivationTime += timeStep */
    if (t > SLEEP_TIME) return n;    /* wantsSleeping(): m_deactivationTime > gDeactivationTime */
  }
  return (int)CNT_MAX + 1;           /* never within the counter's range (h < 0.5 ms): the ball stays awake */
}

/* ---- one substep (Bullet internalSingleStepSimulation, STEP_SPEC 4) ------- */
/* The warm table of one world (STEP_SPEC 4.5a), flat u32 words, the same layout in HBM (include/sslsim_batch.h):
 *   [0, N)        sig[b]    byte k = 1 + id of the static solid in slot k of body b (0 = slot empty), k = 0..3; only the
 *                           ceiling (id 0) and the goal boxes (ids 5..10) live in slots
 *   [N, 5N)       on[k][b]  accumulated normal impulse of that row at the end of the last substep (f32 bits)
 *   [5N, 9N)      of[k][b]  accumulated friction impulse
 *   [9N, 11N)     gn[b], gf[b]     ground row of body b (0 = none)
 *   [11N, 13N)    wxn[b], wxf[b]   its row against a field wall of the x axis (solids 1, 2: a body cannot leave one
 *                                  and reach the other within a substep, so the axis is the key)
 *   [13N, 15N)    wyn[b], wyf[b]   the same for the y axis (solids 3, 4)
 *   [15N]         np        number of pair entries
 *   [15N+1 ...)   pab[cap], pn[cap], pf[cap]   pair rows: a | b << 8, normal and friction impulse
 * N = R + 1, cap = 32 for N <= 16, else 64.  Slot order within a body and entry order within the pair list carry no
 * meaning (the oracle writes canonical order, the kernel arrival order); an entry whose two impulses are zero is the
 * same as no entry; an all-zero table is the empty table. */
int orc_warm_words(int n_robots) {
  const int N = n_robots + 1, cap = N <= 16 ? 32 : MAX_SERIAL_ROWS;
  return 15 * N + 1 + 3 * cap;
}
static float warm_f(const uint32_t *w) { float f; memcpy(&f, w, 4); return f; }
static void warm_set(uint32_t *w, float f) { memcpy(w, &f, 4); }
/* impulses the row's contact point ended the last substep with (0, 0 if it had no row) */
static void warm_lookup(const uint32_t *warm, int N, int cap, const row_t *r, float *pn, float *pf) {
  *pn = 0.0f; *pf = 0.0f;
  if (r->b >= 0) {
    const uint32_t *pt = warm + 15 * N;
    const int np = (int)pt[0];
    const uint32_t key = (uint32_t)(r->a | (r->b << 8));
    for (int c = 0; c < np && c < cap; c++)
      if (pt[1 + c] == key) { *pn = warm_f(&pt[1 + cap + c]); *pf = warm_f(&pt[1 + 2 * cap + c]); return; }
  } else if (r->sid == 11) {
    *pn = warm_f(&warm[9 * N + r->a]); *pf = warm_f(&warm[10 * N + r->a]);
  } else if (r->sid >= 1 && r->sid <= 4) {
    const int o = r->sid <= 2 ? 11 : 13;
    *pn = warm_f(&warm[o * N + r->a]); *pf = warm_f(&warm[(o + 1) * N + r->a]);
  } else {
    const uint32_t sig = warm[r->a];
    for (int k = 0; k < 4; k++)
      if (((sig >> (8 * k)) & 0xffu) == (uint32_t)(r->sid + 1)) {
        *pn = warm_f(&warm[N + k * N + r->a]); *pf = warm_f(&warm[5 * N + k * N + r->a]); return;
      }
  }
}

static void substep(const field_derived *d, const orc_params *p, int R, orc_vec4 *pos,
                    orc_vec4 *vel, uint32_t *aux, const orc_cmd *cmd, float h, int ball_gravity,
                    int sleep_substeps, uint32_t *warm) {
  const int N = R + 1, B = R;
  const float inv_h = 1.0f / h;
  const int spin_mode = (p->flags & ORC_F_BALL_SPIN) != 0;
  float ex[ORC_MAX_BODIES], ey[ORC_MAX_BODIES], ez[ORC_MAX_BODIES], ew[ORC_MAX_BODIES];
  uint32_t cflags[ORC_MAX_BODIES];
  for (int b = 0; b < N; b++) { ex[b] = ey[b] = ez[b] = ew[b] = 0.0f; cflags[b] = 0; }

  /* 4.1 actuation (new functionality; command schema reference protos/grSim_Commands.proto) */
  if (cmd) {
    int kicker = -1, dribbler = -1;
    float kvx = 0, kvy = 0, kvz = 0, dax = 0, day = 0;
    const float bx = pos[B].x, by = pos[B].y, bz = pos[B].z;
    const float bvx = vel[B].x, bvy = vel[B].y;
    for (int r = 0; r < R; r++) {
      const orc_cmd *c = &cmd[r];
      cflags[r] = c->flags;
      float s, co;
      orc_sincos(pos[r].w, &s, &co);
      float dx = bx - pos[r].x, dy = by - pos[r].y;
      float fwd = dot2(co, s, dx, dy);
      float lat = nmad(s, dx, co * dy);
      int grounded = (pos[r].z - WHEEL_R) <= 0.005f;
      int low = (bz - BALL_R) <= p->kick_max_height;
      int in_lat = fabsf(lat) <= p->kick_halfwidth;
      if (grounded && low && in_lat && fwd > 0.0f) {
        if (c->kickspeedx > 0.0f && fwd <= ROBOT_R + BALL_R + p->kick_reach) {
          kicker = r; /* highest index wins */
          kvx = mad(co, c->kickspeedx, vel[r].x);
          kvy = mad(s, c->kickspeedx, vel[r].y);
          kvz = c->kickspeedz;
        }
        if ((c->flags & ORC_CMD_SPINNER) && fwd <= ROBOT_R + BALL_R + p->dribble_reach) {
          dribbler = r;
          float tx = mad(ROBOT_R + BALL_R, co, pos[r].x), ty = mad(ROBOT_R + BALL_R, s, pos[r].y);
          float ax = mad(p->dribble_cd, vel[r].x - bvx, p->dribble_kd * (tx - bx));
          float ay = mad(p->dribble_cd, vel[r].y - bvy, p->dribble_kd * (ty - by));
          float a2 = dot2(ax, ay, ax, ay);
          if (a2 > p->dribble_amax * p->dribble_amax) {
            float k = p->dribble_amax / sqrtf(a2);
            ax = ax * k; ay = ay * k;
          }
          dax = ax; day = ay;
        }
      }
      if ((c->flags & ORC_CMD_DRIVE) && grounded) {
        float u[4];
        if (c->flags & ORC_CMD_WHEELS) {
          for (int i = 0; i < 4; i++) u[i] = c->wheel[i] * WHEEL_R;
        } else {
          for (int i = 0; i < 4; i++)
            u[i] = mad(WHEEL_RC, c->wheel[2], nmad(WSIN[i], c->wheel[0], WCOS[i] * c->wheel[1]));
        }
        float vf = dot2(co, s, vel[r].x, vel[r].y);
        float vl = nmad(s, vel[r].x, co * vel[r].y);
        float ff = 0.0f, fl = 0.0f, tq = 0.0f;
        for (int i = 0; i < 4; i++) {
          float act_i = mad(WHEEL_RC, vel[r].w, nmad(WSIN[i], vf, WCOS[i] * vl));
          float fi = clampf(p->wheel_kp * (u[i] - act_i), -p->wheel_fmax, p->wheel_fmax);
          ff = nmad(WSIN[i], fi, ff);
          fl = mad(WCOS[i], fi, fl);
          tq = tq + fi;
        }
        float k = h / ROBOT_M;
        ex[r] = nmad(s, fl, co * ff) * k;
        ey[r] = mad(co, fl, s * ff) * k;
        ew[r] = (WHEEL_RC * tq) * (h / p->robot_yaw_inertia);
      }
    }
    if (kicker >= 0) {
      vel[B].x = kvx; vel[B].y = kvy; vel[B].z = kvz;
      float s2 = dot3(kvx, kvy, kvz, kvx, kvy, kvz);
      memcpy(&aux[ORC_AUX_PEAK2], &s2, 4);
      aux[ORC_AUX_STATUS] = (aux[ORC_AUX_STATUS] & ~0xFFu) | (uint32_t)kicker | (1u << 15);
      aux[ORC_AUX_OUTS] += 1u << 16;
    }
    if (dribbler >= 0 && kicker < 0) { ex[B] = dax * h; ey[B] = day * h; }
    /* a kicker or dribbler acting on the ball wakes it, like btRigidBody::activate() */
    if (kicker >= 0 || dribbler >= 0) aux[ORC_AUX_STATUS] &= (1u << ACT_SHIFT) - 1u;
  }
  /* 4.2 gravity as an external impulse (Bullet: totalForce * invMass * h).  applyGravity runs once per
   * stepSimulation and skips a body that is asleep at that moment: ball_gravity is latched by orc_step. */
  const float gh = GRAV * h;
  for (int b = 0; b < N; b++) if (b != B || ball_gravity) ez[b] = ez[b] - gh;

  /* solver velocities start at v + external impulse */
  float vx[ORC_MAX_BODIES], vy[ORC_MAX_BODIES], vz[ORC_MAX_BODIES];
  float qx[ORC_MAX_BODIES], qy[ORC_MAX_BODIES], qz[ORC_MAX_BODIES]; /* push velocity */
  float invm[ORC_MAX_BODIES];
  for (int b = 0; b < N; b++) {
    vx[b] = vel[b].x + ex[b]; vy[b] = vel[b].y + ey[b]; vz[b] = vel[b].z + ez[b];
    qx[b] = qy[b] = qz[b] = 0.0f;
    invm[b] = b == B ? 1.0f / BALL_M : 1.0f / ROBOT_M;
  }

  /* the ball's contact margin: Bullet's breaking threshold + the distance it may travel this substep, which stands
   * in for Bullet's predictive (CCD) contact on the ball (reference :58-59) */
  float ball_margin;
  {
    float s2 = dot3(vx[B], vy[B], vz[B], vx[B], vy[B], vz[B]);
    ball_margin = mad(h, sqrtf(s2), MARGIN);
  }

  /* 4.0 simulation islands (Bullet btSimulationIslandManager::buildIslands, restated for a scene whose only
   * body that may sleep is the ball): the ball shares an island with a robot iff their broadphase boxes overlap
   * -- closed form [DEV]: the ball-robot pair passes the pair broadphase, i.e. the horizontal distance of the ball
   * centre from the robot axis is <= ROBOT_R + BALL_R + ball margin (the distance inside which a contact row can
   * exist at all).  An island with a robot never sleeps and wakes a sleeping ball (ISLAND_SLEEPING ->
   * WANTS_DEACTIVATION, time 0); a ball alone sleeps as soon as it wants to. */
  uint32_t act = (aux[ORC_AUX_STATUS] >> ACT_SHIFT) & 3u, cnt = aux[ORC_AUX_STATUS] >> CNT_SHIFT;
  if (act != ACT_ACTIVE) {
    int near = 0;
    const float lim = ROBOT_R + BALL_R + ball_margin;
    for (int r = 0; r < R; r++) {
      float dx = pos[B].x - pos[r].x, dy = pos[B].y - pos[r].y;
      float d2 = dot2(dx, dy, dx, dy);
      near |= d2 <= lim * lim;
    }
    if (near) { if (act == ACT_ASLEEP) { act = ACT_WANTS; cnt = 0; } }
    else act = ACT_ASLEEP;
  }
  const int asleep = act == ACT_ASLEEP; /* not solved, not integrated (btRigidBody::isActive() is false) */
  if (asleep) { /* no external impulse on a sleeping body; no predictive contact either */
    vx[B] = vel[B].x; vy[B] = vel[B].y; vz[B] = vel[B].z;
    ball_margin = MARGIN;
  }

  /* 4.3 contact generation on substep-start positions */
  gen_ctx g;
  g.n = N; g.R = R; g.h = h; g.inv_h = inv_h; g.pos = pos; g.vel = vel; g.spin_mode = spin_mode;
  g.cflags = cmd ? cflags : 0;
  g.hvx = vx; g.hvy = vy; g.hvz = vz;
  g.rest_thresh = p->restitution_threshold;
  g.ball_margin = ball_margin;
  row_t rows[MAX_SERIAL_ROWS];
  row_t ground[ORC_MAX_BODIES];
  int has_ground[ORC_MAX_BODIES];
  const int cap = N <= 16 ? 32 : MAX_SERIAL_ROWS;
  int nrows = 0, overflow = 0;
  for (int i = 0; i < R; i++)
    for (int j = i + 1; j < R; j++) {
      row_t r;
      if (gen_robot_robot(&g, i, j, &r)) { if (nrows < cap) rows[nrows++] = r; else overflow = 1; }
    }
  for (int i = 0; i < R; i++) {
    row_t r;
    if (gen_ball_robot(&g, i, &r)) { if (nrows < cap) rows[nrows++] = r; else overflow = 1; }
  }
  for (int b = 0; b < N; b++) {
    /* at most MAX_BODY_STATICS wall/goal rows per body, first ones in canonical order (STEP_SPEC 4.3) */
    row_t tmp[MAX_BODY_STATICS];
    if (b == B && asleep) continue; /* a sleeping ball's manifolds are not handed to the solver */
    int c = gen_statics(&g, d, b, tmp, MAX_BODY_STATICS);
    if (c > MAX_BODY_STATICS) { overflow = 1; c = MAX_BODY_STATICS; }
    for (int k = 0; k < c; k++) { if (nrows < cap) rows[nrows++] = tmp[k]; else overflow = 1; }
  }
  int nground = 0;
  for (int b = 0; b < N; b++) {
    has_ground[b] = (b == B && asleep) ? 0 : gen_ground(&g, b, &ground[b]);
    nground += has_ground[b];
  }
  if (overflow) aux[ORC_AUX_STATUS] |= 1u << 16;
  aux[ORC_AUX_CONTACTS] += (uint32_t)(nrows + nground);

  /* 4.5a warm starting (Bullet 2.82 setupContactConstraint / setFrictionConstraintImpulse with SOLVER_USE_WARMSTARTING,
   * btContactSolverInfo::m_warmstartingFactor = 0.85): a contact point that persisted from the last substep -- here: a
   * row whose key (bodies, static solid) had a row then -- starts at factor x the impulses it ended that substep with,
   * and the solver velocities receive those impulses at set-up, row by row in solver order, normal before friction
   * (the friction impulse along the row's NEW direction, as Bullet does without friction-direction caching). */
  float wsx = pos[B].w, wsy = vel[B].w; /* ball spin as changed by a warm-started spin friction row */
  if (warm && p->warmstart_factor != 0.0f) {
    const float wf = p->warmstart_factor;
    const float ball_im = 1.0f / BALL_M, sg = (2.5f * ball_im) / BALL_R;
    /* Bullet's refreshContactPoints drops a manifold point whose two anchors have come apart by more than the contact
     * breaking threshold (0.02) along the normal or across it; the narrowphase then adds a new point, which starts
     * cold.  Here: the row is beyond the threshold now (only the ball's speed-dependent margin reaches further), or
     * the two bodies slid by more than the threshold during the last substep (h x lateral relative velocity). */
    const float vbreak = MARGIN * inv_h, l2_break = vbreak * vbreak;
    for (int k = 0; k < nrows + N; k++) {
      row_t *r;
      if (k < nrows) r = &rows[k];
      else { if (!has_ground[k - nrows]) continue; r = &ground[k - nrows]; }
      float pn, pf;
      warm_lookup(warm, N, cap, r, &pn, &pf);
      if (!(r->dist <= MARGIN && r->l2 <= l2_break)) { pn = 0.0f; pf = 0.0f; } /* the point did not persist */
      const int a = r->a, b = r->b;
      const float ima = invm[a], imb = b >= 0 ? invm[b] : 0.0f;
      r->acc_n = wf * pn;
      if (r->acc_n != 0.0f) {
        float ka = r->acc_n * ima;
        vx[a] = mad(r->nx, ka, vx[a]); vy[a] = mad(r->ny, ka, vy[a]); vz[a] = mad(r->nz, ka, vz[a]);
        if (b >= 0) { float kb = r->acc_n * imb; vx[b] = nmad(r->nx, kb, vx[b]); vy[b] = nmad(r->ny, kb, vy[b]); vz[b] = nmad(r->nz, kb, vz[b]); }
      }
      if (r->mu > 0.0f) {
        r->acc_f = wf * pf;
        if (r->acc_f != 0.0f) {
          float ka = r->acc_f * ima;
          vx[a] = mad(r->tx, ka, vx[a]); vy[a] = mad(r->ty, ka, vy[a]); vz[a] = mad(r->tz, ka, vz[a]);
          if (b >= 0) { float kb = r->acc_f * imb; vx[b] = nmad(r->tx, kb, vx[b]); vy[b] = nmad(r->ty, kb, vy[b]); vz[b] = nmad(r->tz, kb, vz[b]); }
          if (r->spin) { float kw = r->acc_f * sg; wsx = mad(r->ty, kw, wsx); wsy = nmad(r->tx, kw, wsy); }
        }
      }
    }
  }

  /* 4.4 split-impulse pass (Bullet solveGroupCacheFriendlySplitImpulseIterations) */
  int deep = 0;
  for (int k = 0; k < nrows; k++) deep |= rows[k].push_target > 0.0f;
  for (int b = 0; b < N; b++) deep |= has_ground[b] && ground[b].push_target > 0.0f;
  if (deep) {
    for (int it = 0; it < p->solver_iterations; it++) {
      for (int k = 0; k < nrows + N; k++) {
        row_t *r;
        if (k < nrows) r = &rows[k];
        else { if (!has_ground[k - nrows]) continue; r = &ground[k - nrows]; }
        if (!(r->push_target > 0.0f)) continue;
        int a = r->a, b = r->b;
        float ima = invm[a], imb = b >= 0 ? invm[b] : 0.0f;
        float meff = 1.0f / (ima + imb);
        float rx = qx[a], ry = qy[a], rz = qz[a];
        if (b >= 0) { rx = rx - qx[b]; ry = ry - qy[b]; rz = rz - qz[b]; }
        float vn = dot3(r->nx, r->ny, r->nz, rx, ry, rz);
        float dl = (r->push_target - vn) * meff;
        float sum = r->acc_p + dl;
        if (sum < 0.0f) { dl = -r->acc_p; sum = 0.0f; }
        r->acc_p = sum;
        float ka = dl * ima;
        qx[a] = mad(r->nx, ka, qx[a]); qy[a] = mad(r->ny, ka, qy[a]); qz[a] = mad(r->nz, ka, qz[a]);
        if (b >= 0) {
          float kb = dl * imb;
          qx[b] = nmad(r->nx, kb, qx[b]); qy[b] = nmad(r->ny, kb, qy[b]); qz[b] = nmad(r->nz, kb, qz[b]);
        }
      }
    }
  }

  /* 4.5 velocity iterations: all normal rows, then all friction rows (Bullet
   * solveSingleIteration without SOLVER_INTERLEAVE_CONTACT_AND_FRICTION) */
  float wx = wsx, wy = wsy; /* ball spin (omega_x, omega_y) */
  const float ball_invm = 1.0f / BALL_M;
  const float spin_meff = 1.0f / (ball_invm + 2.5f * ball_invm);
  const float spin_gain = (2.5f * ball_invm) / BALL_R; /* r / I */
  for (int it = 0; it < p->solver_iterations; it++) {
    for (int k = 0; k < nrows + N; k++) {
      row_t *r;
      if (k < nrows) r = &rows[k];
      else { if (!has_ground[k - nrows]) continue; r = &ground[k - nrows]; }
      int a = r->a, b = r->b;
      float ima = invm[a], imb = b >= 0 ? invm[b] : 0.0f;
      float meff = 1.0f / (ima + imb);
      float rx = vx[a], ry = vy[a], rz = vz[a];
      if (b >= 0) { rx = rx - vx[b]; ry = ry - vy[b]; rz = rz - vz[b]; }
      float vn = dot3(r->nx, r->ny, r->nz, rx, ry, rz);
      float dl = (r->target - vn) * meff;
      float sum = r->acc_n + dl;
      if (sum < 0.0f) { dl = -r->acc_n; sum = 0.0f; }
      r->acc_n = sum;
      float ka = dl * ima;
      vx[a] = mad(r->nx, ka, vx[a]); vy[a] = mad(r->ny, ka, vy[a]); vz[a] = mad(r->nz, ka, vz[a]);
      if (b >= 0) {
        float kb = dl * imb;
        vx[b] = nmad(r->nx, kb, vx[b]); vy[b] = nmad(r->ny, kb, vy[b]); vz[b] = nmad(r->nz, kb, vz[b]);
      }
    }
    for (int k = 0; k < nrows + N; k++) {
      row_t *r;
      if (k < nrows) r = &rows[k];
      else { if (!has_ground[k - nrows]) continue; r = &ground[k - nrows]; }
      if (!(r->mu > 0.0f) || !(r->acc_n > 0.0f)) continue; /* Bullet: friction only if totalImpulse > 0 */
      int a = r->a, b = r->b;
      float ima = invm[a], imb = b >= 0 ? invm[b] : 0.0f;
      float rx = vx[a], ry = vy[a], rz = vz[a];
      if (b >= 0) { rx = rx - vx[b]; ry = ry - vy[b]; rz = rz - vz[b]; }
      float meff;
      if (r->spin) {
        rx = nmad(BALL_R, wy, rx);
        ry = mad(BALL_R, wx, ry);
        meff = spin_meff;
      } else {
        meff = 1.0f / (ima + imb);
      }
      float vt = dot3(r->tx, r->ty, r->tz, rx, ry, rz);
      float dl = (0.0f - vt) * meff;
      float lim = r->mu * r->acc_n;
      float sum = r->acc_f + dl;
      if (sum > lim) { dl = lim - r->acc_f; sum = lim; }
      else if (sum < -lim) { dl = -lim - r->acc_f; sum = -lim; }
      r->acc_f = sum;
      float ka = dl * ima;
      vx[a] = mad(r->tx, ka, vx[a]); vy[a] = mad(r->ty, ka, vy[a]); vz[a] = mad(r->tz, ka, vz[a]);
      if (b >= 0) {
        float kb = dl * imb;
        vx[b] = nmad(r->tx, kb, vx[b]); vy[b] = nmad(r->ty, kb, vy[b]); vz[b] = nmad(r->tz, kb, vz[b]);
      }
      if (r->spin) {
        float kw = dl * spin_gain;
        wx = mad(r->ty, kw, wx);
        wy = nmad(r->tx, kw, wy);
      }
    }
  }

  if (warm && p->warmstart_factor != 0.0f) { /* what the contact points carry into the next substep (Bullet solveGroupCacheFriendlyFinish); with warm starting off the table is left alone */
    uint32_t *pt = warm + 15 * N;
    int np = 0;
    uint32_t wall_seen[ORC_MAX_BODIES];
    for (int b = 0; b < 15 * N; b++) warm[b] = 0;
    for (int b = 0; b < N; b++) wall_seen[b] = 0;
    for (int k = 0; k < nrows + N; k++) {
      const row_t *r;
      if (k < nrows) r = &rows[k];
      else { if (!has_ground[k - nrows]) continue; r = &ground[k - nrows]; }
      if (r->b >= 0) {
        pt[1 + np] = (uint32_t)(r->a | (r->b << 8));
        warm_set(&pt[1 + cap + np], r->acc_n); warm_set(&pt[1 + 2 * cap + np], r->acc_f);
        np++;
      } else if (r->sid == 11) {
        warm_set(&warm[9 * N + r->a], r->acc_n); warm_set(&warm[10 * N + r->a], r->acc_f);
      } else if (r->sid >= 1 && r->sid <= 4) {
        const int o = r->sid <= 2 ? 11 : 13;
        const uint32_t bit = r->sid <= 2 ? 1u : 2u;
        if (!(wall_seen[r->a] & bit)) { /* both walls of an axis at once (a field narrower than a margin): the first */
          wall_seen[r->a] |= bit;
          warm_set(&warm[o * N + r->a], r->acc_n); warm_set(&warm[(o + 1) * N + r->a], r->acc_f);
        }
      } else {
        int slot = 0;
        while (slot < 4 && ((warm[r->a] >> (8 * slot)) & 0xffu)) slot++;
        if (slot < 4) {
          warm[r->a] |= (uint32_t)(r->sid + 1) << (8 * slot);
          warm_set(&warm[N + slot * N + r->a], r->acc_n); warm_set(&warm[5 * N + slot * N + r->a], r->acc_f);
        }
      }
    }
    pt[0] = (uint32_t)np;
  }

  /* 4.6 touches, rolling resistance, peak speed */
  uint32_t touch = 0;
  int last = -1;
  for (int k = 0; k < nrows; k++)
    if (rows[k].a == B && rows[k].b >= 0 && rows[k].acc_n > 0.0f) { touch |= 1u << rows[k].b; last = rows[k].b; }
  if (last >= 0) aux[ORC_AUX_STATUS] = (aux[ORC_AUX_STATUS] & ~0xFFu) | (uint32_t)last;
  aux[ORC_AUX_TOUCH] |= touch;
  if (spin_mode && has_ground[B] && ground[B].acc_n > 0.0f) {
    float s2 = dot2(vx[B], vy[B], vx[B], vy[B]);
    if (s2 > 0.0f) {
      float sp = sqrtf(s2);
      float dec = (p->mu_roll * GRAV) * h;
      float k = sp > dec ? (sp - dec) / sp : 0.0f;
      vx[B] = vx[B] * k; vy[B] = vy[B] * k; wx = wx * k; wy = wy * k;
    }
  }
  {
    float s2 = dot3(vx[B], vy[B], vz[B], vx[B], vy[B], vz[B]);
    float pk; memcpy(&pk, &aux[ORC_AUX_PEAK2], 4);
    if (s2 > pk) memcpy(&aux[ORC_AUX_PEAK2], &s2, 4);
  }

  /* 4.7 write back and integrate (semi-implicit Euler; split-impulse push first) */
  for (int b = 0; b < N; b++) {
    if (b == B && asleep) continue; /* integrateTransforms skips a sleeping body */
    vel[b].x = vx[b]; vel[b].y = vy[b]; vel[b].z = vz[b];
    if (deep) {
      pos[b].x = mad(qx[b], h, pos[b].x); pos[b].y = mad(qy[b], h, pos[b].y); pos[b].z = mad(qz[b], h, pos[b].z);
    }
    pos[b].x = mad(vx[b], h, pos[b].x); pos[b].y = mad(vy[b], h, pos[b].y); pos[b].z = mad(vz[b], h, pos[b].z);
    if (b != B) {
      float wz = vel[b].w + ew[b];
      vel[b].w = wz;
      float yaw = mad(wz, h, pos[b].w);
      if (yaw >= TWO_PI) yaw = yaw - TWO_PI;
      else if (yaw <= -TWO_PI) yaw = yaw + TWO_PI;
      pos[b].w = yaw;
    }
  }
  if (!asleep) { pos[B].w = wx; vel[B].w = wy; }

  /* 4.8 ball activation (Bullet btDiscreteDynamicsWorld::updateActivationState: btRigidBody::updateDeactivation
   * then wantsSleeping).  Angular velocity is the spin state, identically zero in reference mode. */
  if (asleep) {
    vel[B].x = 0.0f; vel[B].y = 0.0f; vel[B].z = 0.0f; /* an ISLAND_SLEEPING body gets its velocities zeroed */
    if (spin_mode) { pos[B].w = 0.0f; vel[B].w = 0.0f; }
  } else {
    float v2 = dot3(vel[B].x, vel[B].y, vel[B].z, vel[B].x, vel[B].y, vel[B].z);
    float w2 = spin_mode ? dot2(pos[B].w, vel[B].w, pos[B].w, vel[B].w) : 0.0f;
    if (v2 < SLEEP_LIN2 && w2 < SLEEP_ANG2) { if (cnt < CNT_MAX) cnt++; }
    else { cnt = 0; act = ACT_ACTIVE; }
    int wants = act == ACT_WANTS || (int)cnt >= sleep_substeps;
    act = wants ? ACT_WANTS : ACT_ACTIVE;
  }
  aux[ORC_AUX_STATUS] = (aux[ORC_AUX_STATUS] & ((1u << ACT_SHIFT) - 1u)) | (act << ACT_SHIFT) | (cnt << CNT_SHIFT);
}

/* ---- step = substeps + events (STEP_SPEC 5) -------------------------------- */
static void step_events(const field_derived *d, int R, const orc_vec4 *pos, uint32_t *aux) {
  const int B = R;
  float x = pos[B].x, y = pos[B].y, z = pos[B].z;
  int mouth = fabsf(y) < d->goal_half_w && z < d->goal_height;
  int in_pos = mouth && x > d->half_len && x < d->goal_back;
  int in_neg = mouth && x < -d->half_len && x > -d->goal_back;
  int out = (fabsf(x) > d->half_len || fabsf(y) > d->half_wid) && !in_pos && !in_neg;
  uint32_t st = aux[ORC_AUX_STATUS];
  int p_pos = (st >> 8) & 1, p_neg = (st >> 9) & 1, p_out = (st >> 10) & 1;
  uint32_t ev = 0;
  if (in_pos && !p_pos) { ev |= 1u << 12; aux[ORC_AUX_GOALS] += 1u; }        /* blue attacks +x */
  if (in_neg && !p_neg) { ev |= 1u << 13; aux[ORC_AUX_GOALS] += 1u << 16; }  /* yellow attacks -x */
  if (out && !p_out) { ev |= 1u << 14; aux[ORC_AUX_OUTS] += 1u; }
  st = (st & ~(7u << 8)) | ((uint32_t)in_pos << 8) | ((uint32_t)in_neg << 9) | ((uint32_t)out << 10);
  st = (st & ~(7u << 12)) | ev; /* bit15 (kicked) was set during the substeps and is kept */
  int bad = 0;
  for (int b = 0; b <= R; b++)
    bad |= !(pos[b].x - pos[b].x == 0.0f) || !(pos[b].y - pos[b].y == 0.0f) || !(pos[b].z - pos[b].z == 0.0f);
  if (bad) st |= 1u << 17;
  aux[ORC_AUX_STATUS] = st;
  aux[ORC_AUX_TOUCHES] += (uint32_t)__builtin_popcount(aux[ORC_AUX_TOUCH]);
}

void orc_step(const orc_field *f, const orc_params *p, int n_robots, orc_vec4 *pos, orc_vec4 *vel,
              uint32_t *aux, const orc_cmd *cmd, int n_steps, int substeps, float h) {
  orc_step_warm(f, p, n_robots, pos, vel, aux, cmd, n_steps, substeps, h, 0);
}

void orc_step_warm(const orc_field *f, const orc_params *p, int n_robots, orc_vec4 *pos, orc_vec4 *vel,
                   uint32_t *aux, const orc_cmd *cmd, int n_steps, int substeps, float h, uint32_t *warm) {
  field_derived d;
  derive_field(f, &d);
  const int sleep_substeps = orc_sleep_substeps(h);
  for (int s = 0; s < n_steps; s++) {
    aux[ORC_AUX_TOUCH] = 0;
    aux[ORC_AUX_STATUS] &= ~(1u << 15);
    /* Bullet stepSimulation: applyGravity() once, for the bodies that are active now (STEP_SPEC 3) */
    const int ball_gravity = ((aux[ORC_AUX_STATUS] >> ACT_SHIFT) & 3u) != ACT_ASLEEP;
    for (int k = 0; k < substeps; k++) substep(&d, p, n_robots, pos, vel, aux, cmd, h, ball_gravity, sleep_substeps, warm);
    step_events(&d, n_robots, pos, aux);
  }
}

/* ---- referee restart placement (STEP_SPEC 7) ------------------------------------ */
int orc_referee(const orc_field *f, int R, orc_vec4 *pos, orc_vec4 *vel, uint32_t *aux) {
  const uint32_t st = aux[ORC_AUX_STATUS];
  const int goal = (st >> 12) & 3, out = (st >> 14) & 1;
  if (!goal && !out) return 0;
  const int B = R;
  float x = 0.0f, y = 0.0f;
  if (!goal) {
    const float lx = f->field_length / 2 - 0.1f, ly = f->field_width / 2 - 0.1f;
    x = clampf(pos[B].x, -lx, lx);
    y = clampf(pos[B].y, -ly, ly);
  }
  pos[B].x = x; pos[B].y = y; pos[B].z = BALL_R; pos[B].w = 0.0f;
  vel[B].x = 0.0f; vel[B].y = 0.0f; vel[B].z = 0.0f; vel[B].w = 0.0f;
  /* not in a goal, not out: the next crossing is a new edge; the placed ball is awake (activation bits 18..31 = 0) */
  aux[ORC_AUX_STATUS] = st & ~(7u << 8) & ((1u << ACT_SHIFT) - 1u);
  aux[ORC_AUX_RESTARTS] += 1u;
  return 1;
}

/* ---- init / commands / gather ---------------------------------------------- */
void orc_init_world(const orc_field *f, int R, uint64_t seed, uint64_t world_id, int mode,
                    orc_vec4 *pos, orc_vec4 *vel, uint32_t *aux) {
  const int B = R;
  for (int b = 0; b <= R; b++) { vel[b].x = vel[b].y = vel[b].z = vel[b].w = 0.0f; }
  for (int k = 0; k < 8; k++) aux[k] = 0;
  aux[ORC_AUX_STATUS] = 0xFFu;
  if (mode == 0) {
    /* reference src/sslsim.cpp:34-42 ranges, :446 drop height; ball :251 */
    float w = f->field_length / 2 - ROBOT_R, hgt = f->field_width / 2 - ROBOT_R;
    for (int r = 0; r < R; r++) {
      float u0 = orc_u01(seed, world_id, (uint64_t)r, 0), u1 = orc_u01(seed, world_id, (uint64_t)r, 1),
            u2 = orc_u01(seed, world_id, (uint64_t)r, 2);
      pos[r].x = -w + u0 * (2.0f * w);
      pos[r].y = -hgt + u1 * (2.0f * hgt);
      pos[r].z = DROP_Z;
      pos[r].w = u2 * TWO_PI;
    }
    pos[B].x = 0.0f; pos[B].y = 0.0f; pos[B].z = DROP_Z; pos[B].w = 0.0f;
  } else {
    /* crowded +x penalty area, resting */
    float x0 = f->field_length / 2 - 1.8f, xw = 1.7f, yw = 1.8f;
    for (int b = 0; b <= R; b++) {
      float x = 0, y = 0;
      for (int t = 0; t < 64; t++) {
        x = x0 + orc_u01(seed, world_id, (uint64_t)b, (uint64_t)(4 * t + 8)) * xw;
        y = -yw + orc_u01(seed, world_id, (uint64_t)b, (uint64_t)(4 * t + 9)) * (2.0f * yw);
        int ok = 1;
        for (int c = 0; c < b; c++) {
          float dx = x - pos[c].x, dy = y - pos[c].y;
          if (dx * dx + dy * dy < 0.19f * 0.19f) ok = 0;
        }
        if (ok) break;
      }
      pos[b].x = x; pos[b].y = y;
      pos[b].z = b == B ? BALL_R : WHEEL_R;
      pos[b].w = b == B ? 0.0f : orc_u01(seed, world_id, (uint64_t)b, 2) * TWO_PI;
    }
  }
}

void orc_gen_commands(int R, uint64_t seed, uint64_t world_id, uint64_t period, int kind, orc_cmd *cmd) {
  for (int r = 0; r < R; r++) {
    uint64_t key = period * 64u + (uint64_t)r;
    for (int i = 0; i < 4; i++) cmd[r].wheel[i] = -40.0f + orc_u01(seed, world_id, key, 100u + (uint64_t)i) * 80.0f;
    cmd[r].flags = ORC_CMD_DRIVE | ORC_CMD_WHEELS;
    cmd[r].kickspeedx = 0.0f; cmd[r].kickspeedz = 0.0f; cmd[r].pad = 0;
    if (kind == 1) {
      cmd[r].flags |= ORC_CMD_SPINNER;
      cmd[r].kickspeedx = orc_u01(seed, world_id, key, 104u) * 6.5f;
      cmd[r].kickspeedz = (r & 1) ? orc_u01(seed, world_id, key, 105u) * 3.0f : 0.0f;
    }
  }
}

/* the orientation the reference's robot_get_pos reports (src/sslsim.cpp:437-443): sign(axis.z) * getAngle() of the
 * quaternion btMatrix3x3::getRotation extracts -- yaw wrapped to (-pi, pi], except that (-pi, -2pi/3) comes back
 * as 2pi - |yaw| (negative-w quaternion of the largest-diagonal branch), and yaw 0 reads -0.0 (identity axis (1,0,0)) */
float orc_readout_yaw(float yaw) {
  float k = rintf(yaw * 0.15915494f);
  float t = yaw - k * TWO_PI;
  if (t < -2.0943952f) t = t + TWO_PI;
  return t == 0.0f ? -0.0f : t;
}

int orc_gather(int R, const orc_vec4 *pos, const int *ids, float *out) {
  int n = 0;
  const orc_vec4 *b = &pos[R];
  out[n++] = 1.0f; out[n++] = b->x; out[n++] = b->y; out[n++] = b->z; out[n++] = b->x; out[n++] = b->y;
  for (int r = 0; r < R; r++) {
    float yaw = orc_readout_yaw(pos[r].w);
    out[n++] = 1.0f; out[n++] = (float)ids[r];
    out[n++] = pos[r].x; out[n++] = pos[r].y; out[n++] = yaw; out[n++] = pos[r].x; out[n++] = pos[r].y;
  }
  return n;
}

/* ---- many worlds, optional threads (timing baseline) ------------------------ */
typedef struct {
  const orc_field *f; const orc_params *p; int R, w0, w1;
  orc_vec4 *pos, *vel; uint32_t *aux; const orc_cmd *cmd; int n_steps, substeps; float h; uint32_t *warm;
} job_t;

static void *job_run(void *arg) {
  job_t *j = (job_t *)arg;
  int N = j->R + 1;
  const size_t ww = (size_t)orc_warm_words(j->R);
  for (int w = j->w0; w < j->w1; w++)
    orc_step_warm(j->f, j->p, j->R, j->pos + (size_t)w * N, j->vel + (size_t)w * N, j->aux + (size_t)w * 8,
                  j->cmd ? j->cmd + (size_t)w * j->R : 0, j->n_steps, j->substeps, j->h,
                  j->warm ? j->warm + (size_t)w * ww : 0);
  return 0;
}

void orc_step_many(const orc_field *f, const orc_params *p, int R, int W, orc_vec4 *pos, orc_vec4 *vel,
                   uint32_t *aux, const orc_cmd *cmd, int n_steps, int substeps, float h, int threads,
                   uint32_t *warm) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  job_t jobs[256];
  pthread_t th[256];
  for (int t = 0; t < threads; t++) {
    job_t j = {f, p, R, (int)((long long)W * t / threads), (int)((long long)W * (t + 1) / threads),
               pos, vel, aux, cmd, n_steps, substeps, h, warm};
    jobs[t] = j;
  }
  if (threads == 1) { job_run(&jobs[0]); return; }
  for (int t = 0; t < threads; t++) pthread_create(&th[t], 0, job_run, &jobs[t]);
  for (int t = 0; t < threads; t++) pthread_join(th[t], 0);
}
