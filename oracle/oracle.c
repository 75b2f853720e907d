/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's hot path (tscircuit/spicey v0.0.13),
 * the fast twin of oracle/spicey_oracle.py.  Used only as the parity checker and as
 * the CPU baseline (tests/, __graft_entry__.smoke(), bench.py cpu_baseline and
 * --impl reference).  Nothing under spicey_b200/ links or loads it.
 *
 * Parity status: PINNED — tests/test_oracle_golden.py checks it against the
 * reference's own golden vectors (tests/golden/) and bit-for-bit against the Python
 * restatement.  Build: oracle/Makefile (gcc -O2 -ffp-contract=off, no -ffast-math, so
 * no operation is fused or re-associated: the arithmetic sequence is the reference's).
 *
 * The reference is TypeScript and no JS runtime exists in this image, so there is no
 * oracle/_ref build.
 *
 * Each function cites the reference file:line it follows.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EPS 1e-15        /* lib/constants/EPS.ts:1 */
#define VT_300K 0.02585  /* lib/constants/physics.ts:1 */

enum { ST_OK = 0, ST_SINGULAR = 1, ST_CDIV = 2, ST_RNONPOS = 3 };

typedef struct { double re, im; } cplx;

/* lib/math/Complex.ts:40-47 — textbook division, throws when |b|^2 < EPS. */
static int cdiv(cplx a, cplx b, cplx *out) {
  double d = b.re * b.re + b.im * b.im;
  if (d < EPS) return ST_CDIV;
  out->re = (a.re * b.re + a.im * b.im) / d;
  out->im = (a.im * b.re - a.re * b.im) / d;
  return ST_OK;
}
static cplx cmul(cplx a, cplx b) { /* Complex.ts:33-38 */
  cplx r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re };
  return r;
}
static cplx csub(cplx a, cplx b) { cplx r = { a.re - b.re, a.im - b.im }; return r; }
static cplx cadd(cplx a, cplx b) { cplx r = { a.re + b.re, a.im + b.im }; return r; }
static double cabs_(cplx a) { return hypot(a.re, a.im); } /* Complex.ts:55-57 */

/* lib/math/solveComplex.ts:4-73.  M is the augmented n x (n+1) matrix, rows reached
 * through `rows` so that a pivot swap exchanges references as the reference does. */
int oracle_solve_complex(int n, cplx *M, cplx **rows, cplx *x) {
  int ld = n + 1;
  for (int i = 0; i < n; i++) rows[i] = M + (size_t)i * ld;
  for (int k = 0; k < n; k++) {
    int imax = k;
    double vmax = cabs_(rows[k][k]);
    for (int i = k + 1; i < n; i++) {        /* :20-28 strict >, first max wins */
      double v = cabs_(rows[i][k]);
      if (v > vmax) { vmax = v; imax = i; }
    }
    if (vmax < EPS) return ST_SINGULAR;       /* :29 */
    if (imax != k) { cplx *t = rows[k]; rows[k] = rows[imax]; rows[imax] = t; }
    cplx *prow = rows[k];
    cplx pivot = prow[k];
    for (int i = k + 1; i < n; i++) {         /* :40-53 */
      cplx *row = rows[i];
      cplx f;
      int st = cdiv(row[k], pivot, &f);
      if (st) return st;
      if (cabs_(f) < EPS) continue;           /* :46 */
      for (int j = k; j <= n; j++) row[j] = csub(row[j], cmul(f, prow[j]));
    }
  }
  for (int i = n - 1; i >= 0; i--) {          /* :56-71 */
    cplx *row = rows[i];
    cplx s = row[n];
    for (int j = i + 1; j < n; j++) s = csub(s, cmul(row[j], x[j]));
    int st = cdiv(s, row[i], &x[i]);
    if (st) return st;
  }
  return ST_OK;
}

/* lib/math/solveReal.ts:3-73 */
int oracle_solve_real(int n, double *M, double **rows, double *x) {
  int ld = n + 1;
  for (int i = 0; i < n; i++) rows[i] = M + (size_t)i * ld;
  for (int k = 0; k < n; k++) {
    int imax = k;
    double vmax = fabs(rows[k][k]);
    for (int i = k + 1; i < n; i++) {
      double v = fabs(rows[i][k]);
      if (v > vmax) { vmax = v; imax = i; }
    }
    if (vmax < EPS) return ST_SINGULAR;       /* :28 */
    if (imax != k) { double *t = rows[k]; rows[k] = rows[imax]; rows[imax] = t; }
    double *prow = rows[k];
    double pivot = prow[k];
    for (int i = k + 1; i < n; i++) {
      double *row = rows[i];
      double f = row[k] / pivot;
      if (fabs(f) < EPS) continue;
      for (int j = k; j <= n; j++) row[j] = row[j] - f * prow[j];
    }
  }
  for (int i = n - 1; i >= 0; i--) {
    double *row = rows[i];
    double s = row[n];
    for (int j = i + 1; j < n; j++) s -= row[j] * x[j];
    x[i] = s / row[i];
  }
  return ST_OK;
}

/* Flat circuit: per-type arrays in netlist order (ParsedCircuit, parseNetlist.ts:83-103).
 * Value arrays are [n_inst][count]; node arrays are shared by all instances. */
typedef struct {
  int32_t nn, nR, nC, nL, nV, nS, nD;
  const int32_t *r_n1, *r_n2; const double *r_val;
  const int32_t *c_n1, *c_n2; const double *c_val;
  const int32_t *l_n1, *l_n2; const double *l_val;
  const int32_t *v_n1, *v_n2; const double *v_dc, *v_acmag, *v_acphase;
  const int32_t *s_n1, *s_n2, *s_cp, *s_cn; const double *s_ron, *s_roff, *s_von, *s_voff;
  const int32_t *d_np, *d_nm; const double *d_is, *d_n;
} ocircuit;

#define MI(n) ((n) - 1) /* NodeIndex.ts:28-31: ground (0) -> -1 */

static void stamp_y_c(cplx *A, int ld, int n1, int n2, cplx Y) { /* stampAdmittanceComplex.ts:4-30 */
  int i1 = MI(n1), i2 = MI(n2);
  if (i1 >= 0) A[i1 * ld + i1] = cadd(A[i1 * ld + i1], Y);
  if (i2 >= 0) A[i2 * ld + i2] = cadd(A[i2 * ld + i2], Y);
  if (i1 >= 0 && i2 >= 0) {
    A[i1 * ld + i2] = csub(A[i1 * ld + i2], Y);
    A[i2 * ld + i1] = csub(A[i2 * ld + i1], Y);
  }
}
static void stamp_y_r(double *A, int ld, int n1, int n2, double Y) { /* stampAdmittanceReal.ts:3-29 */
  int i1 = MI(n1), i2 = MI(n2);
  if (i1 >= 0) A[i1 * ld + i1] += Y;
  if (i2 >= 0) A[i2 * ld + i2] += Y;
  if (i1 >= 0 && i2 >= 0) { A[i1 * ld + i2] -= Y; A[i2 * ld + i1] -= Y; }
}
static void stamp_i_r(double *A, int ld, int np, int nm, double I) { /* stampCurrentReal.ts:3-14 (b = column ld-1) */
  int ip = MI(np), im = MI(nm), n = ld - 1;
  if (ip >= 0) A[ip * ld + n] = A[ip * ld + n] - I;
  if (im >= 0) A[im * ld + n] = A[im * ld + n] + I;
}

static int ind_adm(double f, double L, cplx *Y) { /* simulateAC.ts:47-51 */
  cplx denom = { 0.0, 2 * M_PI * f * L };
  if (cabs_(denom) < EPS) { Y->re = 0; Y->im = 0; return ST_OK; }
  cplx one = { 1, 0 };
  return cdiv(one, denom, Y);
}

/* One AC point: simulateAC.ts:24-60 (build) + :83 (solve) + :85-126 (unpack).
 * x: [nvar] complex; ielem: [nR+nC+nL+nV] complex in R,C,L,V order. */
static int ac_point(const ocircuit *c, int inst, double f, cplx *M, cplx **rows, cplx *x, cplx *ielem) {
  int nvar = c->nn + c->nV, ld = nvar + 1;
  const double twoPi = 2 * M_PI;
  const double *rv = c->r_val + (size_t)inst * c->nR, *cv = c->c_val + (size_t)inst * c->nC;
  const double *lv = c->l_val + (size_t)inst * c->nL;
  const double *vm = c->v_acmag + (size_t)inst * c->nV, *vp = c->v_acphase + (size_t)inst * c->nV;
  memset(M, 0, sizeof(cplx) * nvar * ld);
  for (int i = 0; i < c->nR; i++) {
    if (rv[i] <= 0) return ST_RNONPOS;             /* :37 */
    cplx Y = { 1 / rv[i], 0 };
    stamp_y_c(M, ld, c->r_n1[i], c->r_n2[i], Y);
  }
  for (int i = 0; i < c->nC; i++) {
    cplx Y = { 0, twoPi * f * cv[i] };
    stamp_y_c(M, ld, c->c_n1[i], c->c_n2[i], Y);
  }
  for (int i = 0; i < c->nL; i++) {
    cplx Y; int st = ind_adm(f, lv[i], &Y);
    if (st) return st;
    stamp_y_c(M, ld, c->l_n1[i], c->l_n2[i], Y);
  }
  for (int i = 0; i < c->nV; i++) {                /* stampVoltageSourceComplex.ts:5-35 */
    double ph = ((vp[i]) * M_PI) / 180;            /* Complex.ts:16-19 */
    cplx V = { vm[i] * cos(ph), vm[i] * sin(ph) };
    int i1 = MI(c->v_n1[i]), i2 = MI(c->v_n2[i]), j = c->nn + i;
    cplx one = { 1, 0 };
    if (i1 >= 0) M[i1 * ld + j] = cadd(M[i1 * ld + j], one);
    if (i2 >= 0) M[i2 * ld + j] = csub(M[i2 * ld + j], one);
    if (i1 >= 0) M[j * ld + i1] = cadd(M[j * ld + i1], one);
    if (i2 >= 0) M[j * ld + i2] = csub(M[j * ld + i2], one);
    M[j * ld + nvar] = cadd(M[j * ld + nvar], V);
  }
  int st = oracle_solve_complex(nvar, M, rows, x);
  if (st) return st;
  if (ielem) {
    cplx zero = { 0, 0 };
    int e = 0;
#define VOLT(n) ((n) == 0 ? zero : x[(n) - 1])
    for (int i = 0; i < c->nR; i++, e++) {
      cplx Y = { 1 / rv[i], 0 };
      ielem[e] = cmul(Y, csub(VOLT(c->r_n1[i]), VOLT(c->r_n2[i])));
    }
    for (int i = 0; i < c->nC; i++, e++) {
      cplx Y = { 0, twoPi * f * cv[i] };
      ielem[e] = cmul(Y, csub(VOLT(c->c_n1[i]), VOLT(c->c_n2[i])));
    }
    for (int i = 0; i < c->nL; i++, e++) {
      cplx Y; st = ind_adm(f, lv[i], &Y);
      if (st) return st;
      ielem[e] = cmul(Y, csub(VOLT(c->l_n1[i]), VOLT(c->l_n2[i])));
    }
    for (int i = 0; i < c->nV; i++, e++) ielem[e] = x[c->nn + i];
#undef VOLT
  }
  return ST_OK;
}

typedef struct {
  const ocircuit *c; const double *freqs; int64_t F, lo, hi;
  double *x, *ielem; int32_t *status;
} ac_job;

static void *ac_worker(void *arg) {
  ac_job *j = (ac_job *)arg;
  const ocircuit *c = j->c;
  int nvar = c->nn + c->nV, nel = c->nR + c->nC + c->nL + c->nV;
  cplx *M = (cplx *)malloc(sizeof(cplx) * nvar * (nvar + 1));
  cplx **rows = (cplx **)malloc(sizeof(cplx *) * nvar);
  for (int64_t p = j->lo; p < j->hi; p++) {
    int inst = (int)(p / j->F);
    double f = j->freqs[p % j->F];
    cplx *x = (cplx *)(j->x + (size_t)p * nvar * 2);
    cplx *ie = j->ielem ? (cplx *)(j->ielem + (size_t)p * nel * 2) : NULL;
    j->status[p] = ac_point(c, inst, f, M, rows, x, ie);
  }
  free(M); free(rows);
  return NULL;
}

/* simulateAC.ts:80-127 over points p = inst*F + f; contiguous ranges per thread. */
int oracle_ac(const ocircuit *c, const double *freqs, int64_t F, int64_t n_inst,
              double *x, double *ielem, int32_t *status, int nthreads) {
  int64_t P = F * n_inst;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > P) nthreads = (int)(P > 0 ? P : 1);
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  ac_job *jobs = (ac_job *)malloc(sizeof(ac_job) * nthreads);
  for (int t = 0; t < nthreads; t++) {
    ac_job jb = { c, freqs, F, P * t / nthreads, P * (t + 1) / nthreads, x, ielem, status };
    jobs[t] = jb;
    if (nthreads == 1) ac_worker(&jobs[t]);
    else pthread_create(&th[t], NULL, ac_worker, &jobs[t]);
  }
  if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
  return 0;
}

/* ---- TRAN: simulateTRAN.ts:130-252 for one instance -------------------------------
 * vsrc: [nV][steps+1] pre-sampled waveform(t_k) (host-evaluated; SURVEY H-G), used
 * where v_wave[i] != 0, else dc.  state: vPrev[nC], iPrev[nL], vdPrev[nD], isOn[nS]
 * (as doubles 0/1), read at entry and written back at exit (the reference mutates ckt).
 * out_v: [steps+1][nn]; out_i: [steps+1][nR+nC+nL+nV+nS+nD]; iters: [steps+1] solves. */
static int tran_instance(const ocircuit *c, int inst, double dt, int64_t steps, const double *vsrc,
                         const int32_t *v_wave, double *state, double *out_v, double *out_i,
                         int32_t *iters) {
  int nvar = c->nn + c->nV, ld = nvar + 1;
  int nel = c->nR + c->nC + c->nL + c->nV + c->nS + c->nD;
  const double *rv = c->r_val + (size_t)inst * c->nR, *cv = c->c_val + (size_t)inst * c->nC;
  const double *lv = c->l_val + (size_t)inst * c->nL, *vdc = c->v_dc + (size_t)inst * c->nV;
  const double *ron = c->s_ron + (size_t)inst * c->nS, *roff = c->s_roff + (size_t)inst * c->nS;
  const double *von = c->s_von + (size_t)inst * c->nS, *voff = c->s_voff + (size_t)inst * c->nS;
  const double *dis = c->d_is + (size_t)inst * c->nD, *dn = c->d_n + (size_t)inst * c->nD;
  double *vPrev = state, *iPrev = vPrev + c->nC, *vdPrev = iPrev + c->nL, *isOn = vdPrev + c->nD;
  double *M = (double *)malloc(sizeof(double) * nvar * ld);
  double **rows = (double **)malloc(sizeof(double *) * nvar);
  double *x = (double *)malloc(sizeof(double) * (nvar + 1));
  int rc = ST_OK;
  const double dtc = fmax(dt, EPS);
#define VOLT(n) ((n) == 0 ? 0.0 : x[(n) - 1])
  for (int64_t step = 0; step <= steps && rc == ST_OK; step++) {   /* :146-147 */
    for (int i = 0; i < nvar; i++) x[i] = 0.0;                     /* :149 */
    int it;
    for (it = 0; it < 20; it++) {                                  /* :151 */
      memset(M, 0, sizeof(double) * nvar * ld);
      for (int i = 0; i < c->nR; i++) stamp_y_r(M, ld, c->r_n1[i], c->r_n2[i], 1 / rv[i]);
      for (int i = 0; i < c->nC; i++) {                            /* :41-46 */
        double Gc = cv[i] / dtc;
        stamp_y_r(M, ld, c->c_n1[i], c->c_n2[i], Gc);
        stamp_i_r(M, ld, c->c_n1[i], c->c_n2[i], -Gc * vPrev[i]);
      }
      for (int i = 0; i < c->nL; i++) {                            /* :49-53 */
        double Gl = dtc / lv[i];
        stamp_y_r(M, ld, c->l_n1[i], c->l_n2[i], Gl);
        stamp_i_r(M, ld, c->l_n1[i], c->l_n2[i], iPrev[i]);
      }
      for (int i = 0; i < c->nS; i++) {                            /* :56-63 */
        double Rv = isOn[i] != 0.0 ? ron[i] : roff[i];
        stamp_y_r(M, ld, c->s_n1[i], c->s_n2[i], 1 / fmax(fabs(Rv), EPS));
      }
      for (int i = 0; i < c->nV; i++) {                            /* :66-69 + stampVoltageSourceReal.ts */
        double Vt = v_wave[i] ? vsrc[(size_t)i * (steps + 1) + step] : vdc[i];
        int i1 = MI(c->v_n1[i]), i2 = MI(c->v_n2[i]), j = c->nn + i;
        if (i1 >= 0) M[i1 * ld + j] += 1;
        if (i2 >= 0) M[i2 * ld + j] -= 1;
        if (i1 >= 0) M[j * ld + i1] += 1;
        if (i2 >= 0) M[j * ld + i2] -= 1;
        M[j * ld + nvar] += Vt;
      }
      for (int i = 0; i < c->nD; i++) {                            /* :72-101 */
        double vd = it == 0 ? vdPrev[i] : VOLT(c->d_np[i]) - VOLT(c->d_nm[i]);
        double vth = dn[i] * VT_300K;
        double vlim = vd;
        if (vd > 0.8) vlim = 0.8;
        if (vd < -1.0) vlim = -1.0;
        double e = exp(vlim / vth);
        double id = dis[i] * (e - 1);
        double gd = fmax((dis[i] / vth) * e, 1e-12);
        double ieq = id - gd * vlim;
        stamp_y_r(M, ld, c->d_np[i], c->d_nm[i], gd);
        stamp_i_r(M, ld, c->d_np[i], c->d_nm[i], ieq);
      }
      rc = oracle_solve_real(nvar, M, rows, x);
      if (rc) break;
      int switched = 0;                                            /* :108-128 */
      for (int i = 0; i < c->nS; i++) {
        double vctrl = VOLT(c->s_cp[i]) - VOLT(c->s_cn[i]);
        int on = isOn[i] != 0.0, nxt = on;
        if (on) { if (vctrl < voff[i]) nxt = 0; }
        else if (vctrl > von[i]) nxt = 1;
        if (nxt != on) { isOn[i] = nxt; switched = 1; }
      }
      if (!switched) break;
    }
    if (rc) break;
    if (iters) iters[step] = it < 20 ? it + 1 : 20;
    double *ov = out_v + (size_t)step * c->nn;
    for (int i = 0; i < c->nn; i++) ov[i] = x[i];                  /* :164-171 */
    if (out_i) {
      double *oi = out_i + (size_t)step * nel;
      int e = 0;
      for (int i = 0; i < c->nR; i++, e++) oi[e] = (VOLT(c->r_n1[i]) - VOLT(c->r_n2[i])) / rv[i];
      for (int i = 0; i < c->nC; i++, e++)
        oi[e] = (cv[i] * (VOLT(c->c_n1[i]) - VOLT(c->c_n2[i]) - vPrev[i])) / dtc;
      for (int i = 0; i < c->nL; i++, e++) {
        double Gl = dtc / lv[i];
        oi[e] = Gl * (VOLT(c->l_n1[i]) - VOLT(c->l_n2[i])) + iPrev[i];
      }
      for (int i = 0; i < c->nV; i++, e++) oi[e] = x[c->nn + i];
      for (int i = 0; i < c->nS; i++, e++) {
        double Rv = isOn[i] != 0.0 ? ron[i] : roff[i];
        oi[e] = (VOLT(c->s_n1[i]) - VOLT(c->s_n2[i])) / fmax(fabs(Rv), EPS);
      }
      for (int i = 0; i < c->nD; i++, e++) {                       /* :208-219 unclamped vd */
        double vd = VOLT(c->d_np[i]) - VOLT(c->d_nm[i]);
        oi[e] = dis[i] * (exp(vd / (dn[i] * VT_300K)) - 1);
      }
    }
    for (int i = 0; i < c->nC; i++) vPrev[i] = VOLT(c->c_n1[i]) - VOLT(c->c_n2[i]);   /* :221-225 */
    for (int i = 0; i < c->nL; i++) {                                                /* :226-231 */
      double Gl = dtc / lv[i];
      iPrev[i] = Gl * (VOLT(c->l_n1[i]) - VOLT(c->l_n2[i])) + iPrev[i];
    }
    for (int i = 0; i < c->nD; i++) vdPrev[i] = VOLT(c->d_np[i]) - VOLT(c->d_nm[i]); /* :233-237 */
  }
#undef VOLT
  free(M); free(rows); free(x);
  return rc;
}

typedef struct {
  const ocircuit *c; double dt; int64_t steps; const double *vsrc; const int32_t *v_wave;
  double *state, *out_v, *out_i; int32_t *iters, *status; int64_t lo, hi;
} tran_job;

static void *tran_worker(void *arg) {
  tran_job *j = (tran_job *)arg;
  const ocircuit *c = j->c;
  int nst = c->nC + c->nL + c->nD + c->nS;
  int nel = c->nR + c->nC + c->nL + c->nV + c->nS + c->nD;
  size_t S1 = (size_t)j->steps + 1;
  for (int64_t i = j->lo; i < j->hi; i++)
    j->status[i] = tran_instance(c, (int)i, j->dt, j->steps, j->vsrc, j->v_wave,
                                 j->state + (size_t)i * nst, j->out_v + (size_t)i * S1 * c->nn,
                                 j->out_i ? j->out_i + (size_t)i * S1 * nel : NULL,
                                 j->iters ? j->iters + (size_t)i * S1 : NULL);
  return NULL;
}

/* Batch of independent instances (the caller-side loop the reference lacks).
 * Layouts are instance-major: out_v[inst][step][node], out_i[inst][step][elem]. */
int oracle_tran(const ocircuit *c, double dt, int64_t steps, const double *vsrc, const int32_t *v_wave,
                int64_t n_inst, double *state, double *out_v, double *out_i, int32_t *iters,
                int32_t *status, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n_inst) nthreads = (int)(n_inst > 0 ? n_inst : 1);
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  tran_job *jobs = (tran_job *)malloc(sizeof(tran_job) * nthreads);
  for (int t = 0; t < nthreads; t++) {
    tran_job jb = { c, dt, steps, vsrc, v_wave, state, out_v, out_i, iters, status,
                    n_inst * t / nthreads, n_inst * (t + 1) / nthreads };
    jobs[t] = jb;
    if (nthreads == 1) tran_worker(&jobs[t]);
    else pthread_create(&th[t], NULL, tran_worker, &jobs[t]);
  }
  if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
  return 0;
}

/* Thin wrappers for direct solver tests (interleaved re,im; A row-major n x n). */
int oracle_solve_complex_flat(int n, const double *A, const double *b, double *x) {
  cplx *M = (cplx *)malloc(sizeof(cplx) * n * (n + 1));
  cplx **rows = (cplx **)malloc(sizeof(cplx *) * n);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) { M[i * (n + 1) + j].re = A[2 * (i * n + j)]; M[i * (n + 1) + j].im = A[2 * (i * n + j) + 1]; }
    M[i * (n + 1) + n].re = b[2 * i]; M[i * (n + 1) + n].im = b[2 * i + 1];
  }
  int st = oracle_solve_complex(n, M, rows, (cplx *)x);
  free(M); free(rows);
  return st;
}
int oracle_solve_real_flat(int n, const double *A, const double *b, double *x) {
  double *M = (double *)malloc(sizeof(double) * n * (n + 1));
  double **rows = (double **)malloc(sizeof(double *) * n);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) M[i * (n + 1) + j] = A[i * n + j];
    M[i * (n + 1) + n] = b[i];
  }
  int st = oracle_solve_real(n, M, rows, x);
  free(M); free(rows);
  return st;
}
