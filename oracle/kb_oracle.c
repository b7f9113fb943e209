/*
 * kb_oracle.c -- CPU oracle (plain C, fp64) for the batched configuration-feasibility path.
 *
 * TEST INFRASTRUCTURE ONLY -- see kb_oracle.h.  PARITY UNPINNED for the collision / distance arithmetic
 * (KrisLibrary absent; no golden vectors in the reference).  Pinned by outputs of the reference's own Python
 * code, generated here and committed under tests/golden/: forward kinematics (KinematicsBuilder of
 * Python/klampt/math/autodiff/kinematics_ad.py:407-457, tests/golden/make_reference_fk.py), the SO(3) arithmetic of Floating / BallAndSocket joints (z-y-x FK,
 * geodesic interpolation, angle metric), which tests/test_reference_golden.py checks against outputs of
 * the reference's own Python/klampt/math/so3.py (tests/golden/make_reference_so3.py), and the default pair
 * mask, which is checked against the reference's own WorldCollider (tests/golden/make_reference_mask.py).
 * Deliberately simple: fp64 everywhere, median-split AABB trees with
 * one element per leaf (8 for point clouds), exhaustive minima, no FMA contraction
 * (-ffp-contract=off in the Makefile).
 *
 * Reference anchors (relative to /root/reference):
 *   FK recurrence ............ Python/klampt/math/autodiff/kinematics_ad.py:434-457,
 *                              Cpp/docs/Manual-Modeling.md:94,107-117
 *   joint / driver limits .... Cpp/Planning/RobotCSpace.cpp:610-630, Cpp/Modeling/Robot.cpp:2166-2187
 *   IsFeasible ............... Cpp/Planning/RobotCSpace.cpp:786-823
 *   pair mask ................ Cpp/Planning/PlannerSettings.cpp:16-41
 *   self-collision defaults .. Cpp/docs/Manual-FileTypes.md:212, Cpp/Modeling/Robot.cpp:1274-1313
 *   broad phase .............. Cpp/Planning/PlannerSettings.cpp:241-331
 *   pair query ............... Cpp/Planning/PlannerSettings.cpp:96-115
 *   margin semantics ......... Cpp/docs/Manual-Geometry.md:17, Python/klampt/src/geometry.h:1006-1108
 *   edge checker ............. Cpp/Planning/RobotCSpace.cpp:835-838 (EpsilonEdgeChecker, eps from
 *                              PlannerSettings.cpp:86-87), metric/interp Cpp/Modeling/Interpolate.cpp:10-71,208-343
 *   distance lower bound ..... Cpp/Planning/PlannerSettings.cpp:570-620
 */
#include "kb_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <float.h>
#include <alloca.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ small vector math */
typedef struct { double R[9]; double t[3]; } xf_t;

static inline void v_sub(const double* a, const double* b, double* o) { o[0]=a[0]-b[0]; o[1]=a[1]-b[1]; o[2]=a[2]-b[2]; }
static inline void v_add(const double* a, const double* b, double* o) { o[0]=a[0]+b[0]; o[1]=a[1]+b[1]; o[2]=a[2]+b[2]; }
static inline double v_dot(const double* a, const double* b) { return a[0]*b[0]+a[1]*b[1]+a[2]*b[2]; }
static inline void v_cross(const double* a, const double* b, double* o) {
  o[0]=a[1]*b[2]-a[2]*b[1]; o[1]=a[2]*b[0]-a[0]*b[2]; o[2]=a[0]*b[1]-a[1]*b[0]; }
static inline void v_madd(const double* a, const double* d, double s, double* o) { o[0]=a[0]+s*d[0]; o[1]=a[1]+s*d[1]; o[2]=a[2]+s*d[2]; }
static inline double v_dist2(const double* a, const double* b) { double d[3]; v_sub(a,b,d); return v_dot(d,d); }

static void xf_identity(xf_t* T) { memset(T,0,sizeof(*T)); T->R[0]=T->R[4]=T->R[8]=1.0; }
static void xf_from12(const double* a, xf_t* T) { memcpy(T->R,a,9*sizeof(double)); memcpy(T->t,a+9,3*sizeof(double)); }
static void xf_to12(const xf_t* T, double* a) { memcpy(a,T->R,9*sizeof(double)); memcpy(a+9,T->t,3*sizeof(double)); }
static inline void xf_apply(const xf_t* T, const double* p, double* o) {
  double x=p[0],y=p[1],z=p[2];
  o[0]=T->R[0]*x+T->R[1]*y+T->R[2]*z+T->t[0];
  o[1]=T->R[3]*x+T->R[4]*y+T->R[5]*z+T->t[1];
  o[2]=T->R[6]*x+T->R[7]*y+T->R[8]*z+T->t[2];
}
/* C = A*B */
static void xf_mul(const xf_t* A, const xf_t* B, xf_t* C) {
  xf_t o;
  for (int i=0;i<3;i++) for (int j=0;j<3;j++)
    o.R[3*i+j]=A->R[3*i]*B->R[j]+A->R[3*i+1]*B->R[3+j]+A->R[3*i+2]*B->R[6+j];
  for (int i=0;i<3;i++) o.t[i]=A->R[3*i]*B->t[0]+A->R[3*i+1]*B->t[1]+A->R[3*i+2]*B->t[2]+A->t[i];
  *C=o;
}
/* C = A^-1 * B */
static void xf_mul_inv_a(const xf_t* A, const xf_t* B, xf_t* C) {
  xf_t o; double d[3]; v_sub(B->t,A->t,d);
  for (int i=0;i<3;i++) for (int j=0;j<3;j++)
    o.R[3*i+j]=A->R[i]*B->R[j]+A->R[3+i]*B->R[3+j]+A->R[6+i]*B->R[6+j];
  for (int i=0;i<3;i++) o.t[i]=A->R[i]*d[0]+A->R[3+i]*d[1]+A->R[6+i]*d[2];
  *C=o;
}

/* ------------------------------------------------------------------ geometry containers */
enum { G_EMPTY=0, G_MESH=1, G_CLOUD=2, G_PRIM=3 };
#define CLOUD_LEAF 8

typedef struct { double lo[3], hi[3]; int left, right; int first; int count; } node_t; /* left<0 => leaf [first,first+count) */

typedef struct {
  int kind;
  double margin;
  int nv, nt;           /* mesh */
  double* verts;        /* nv*3 */
  int32_t* tris;        /* nt*3 */
  double* tv;           /* nt*9 expanded triangle vertices, BVH order */
  int np;               /* cloud / prim (sphere): points + radius */
  double* pts;          /* np*3, BVH order */
  double* rad;          /* np (may be all zeros) */
  int* perm;            /* element permutation (BVH order -> original index) */
  node_t* nodes; int nnodes;
  double lo[3], hi[3];  /* local AABB */
  double rmax;          /* largest element radius (0 for meshes) */
  int solid;            /* box primitive: the mesh holds its 12 surface triangles, and the interior counts too */
  double bc[3], bR[9], bh[3];   /* solid box: centre, axes (columns of the row-major bR), half dimensions, local frame */
} geom_t;

typedef struct { int n; int32_t* links; double* scale; double* offset; double dmin, dmax; } driver_t;

struct ko_world {
  geom_t* geoms; int ngeoms, capgeoms;
  int* terrains; int nterr;
  int* objects; xf_t* objT; int nobj;
  int L; int32_t* parents; uint8_t* linktype; double* axis; xf_t* T0; double* qmin; double* qmax;
  int* linkgeom;
  int nj; uint8_t* jtype; int32_t* jlink; int32_t* jbase;
  driver_t* drivers; int ndrv;
  uint8_t* selfcol;     /* L*L upper triangular (i<j) */
  int selfcol_default;
  uint8_t* mask; int nids; int mask_user;
  int finalized;
};

/* ------------------------------------------------------------------ element predicates */
static inline double orient3d(const double* a, const double* b, const double* c, const double* d) {
  double ad[3],bd[3],cd[3],bc[3];
  v_sub(a,d,ad); v_sub(b,d,bd); v_sub(c,d,cd);
  v_cross(bd,cd,bc);
  return v_dot(ad,bc);
}

static inline double orient2d(const double* a, const double* b, const double* c) {
  return (b[0]-a[0])*(c[1]-a[1]) - (b[1]-a[1])*(c[0]-a[0]);
}
static int on_seg2d(const double* a, const double* b, const double* p) {
  return p[0] >= fmin(a[0],b[0]) && p[0] <= fmax(a[0],b[0]) && p[1] >= fmin(a[1],b[1]) && p[1] <= fmax(a[1],b[1]);
}
static int seg_seg_2d(const double* a, const double* b, const double* c, const double* d) {
  double o1=orient2d(a,b,c), o2=orient2d(a,b,d), o3=orient2d(c,d,a), o4=orient2d(c,d,b);
  if (((o1>0&&o2<0)||(o1<0&&o2>0)) && ((o3>0&&o4<0)||(o3<0&&o4>0))) return 1;
  if (o1==0 && on_seg2d(a,b,c)) return 1;
  if (o2==0 && on_seg2d(a,b,d)) return 1;
  if (o3==0 && on_seg2d(c,d,a)) return 1;
  if (o4==0 && on_seg2d(c,d,b)) return 1;
  return 0;
}
static int point_in_tri_2d(const double* p, const double* a, const double* b, const double* c) {
  double o1=orient2d(a,b,p), o2=orient2d(b,c,p), o3=orient2d(c,a,p);
  return (o1>=0&&o2>=0&&o3>=0)||(o1<=0&&o2<=0&&o3<=0);
}
/* coplanar triangles: project on the dominant plane of the normal n */
static int tri_tri_coplanar(const double* A, const double* B, const double* n) {
  int ax=0; double m=fabs(n[0]);
  if (fabs(n[1])>m) { ax=1; m=fabs(n[1]); }
  if (fabs(n[2])>m) { ax=2; }
  int i0=(ax+1)%3, i1=(ax+2)%3;
  double a[3][2], b[3][2];
  for (int k=0;k<3;k++) { a[k][0]=A[3*k+i0]; a[k][1]=A[3*k+i1]; b[k][0]=B[3*k+i0]; b[k][1]=B[3*k+i1]; }
  for (int i=0;i<3;i++) for (int j=0;j<3;j++)
    if (seg_seg_2d(a[i],a[(i+1)%3],b[j],b[(j+1)%3])) return 1;
  if (point_in_tri_2d(a[0],b[0],b[1],b[2])) return 1;
  if (point_in_tri_2d(b[0],a[0],a[1],a[2])) return 1;
  return 0;
}
/* closed segment pq vs closed triangle abc, pq not coplanar with abc as a whole */
static int seg_tri(const double* p, const double* q, const double* a, const double* b, const double* c, double sp, double sq) {
  if ((sp>0&&sq>0)||(sp<0&&sq<0)) return 0;
  if (sp==0 && sq==0) return 0; /* segment inside the plane: caught by the other triangle's edges or coplanar path */
  double s1=orient3d(p,q,a,b), s2=orient3d(p,q,b,c), s3=orient3d(p,q,c,a);
  return (s1>=0&&s2>=0&&s3>=0)||(s1<=0&&s2<=0&&s3<=0);
}

static double seg_seg_dist2(const double* p1, const double* q1, const double* p2, const double* q2);

/* Two closed triangles share a point?  Edge-vs-triangle formulation (any exact tri-tri overlap test is
 * acceptable because the boolean is a geometric fact). */
int ko_tri_tri_intersect(const double A[9], const double B[9]) {
  const double *a0=A,*a1=A+3,*a2=A+6,*b0=B,*b1=B+3,*b2=B+6;
  double da[3], db[3];
  da[0]=orient3d(b0,b1,b2,a0); da[1]=orient3d(b0,b1,b2,a1); da[2]=orient3d(b0,b1,b2,a2);
  if ((da[0]>0&&da[1]>0&&da[2]>0)||(da[0]<0&&da[1]<0&&da[2]<0)) return 0;
  db[0]=orient3d(a0,a1,a2,b0); db[1]=orient3d(a0,a1,a2,b1); db[2]=orient3d(a0,a1,a2,b2);
  if ((db[0]>0&&db[1]>0&&db[2]>0)||(db[0]<0&&db[1]<0&&db[2]<0)) return 0;
  /* Coplanar only if EACH triangle lies in the other's plane.  A zero-area triangle (repeated or collinear vertices) makes every
   * orient3d against it vanish without the other triangle being anywhere near its line: that case goes on to the edge tests,
   * where its edges are tested against the proper triangle like any segment. */
  if (da[0]==0&&da[1]==0&&da[2]==0 && db[0]==0&&db[1]==0&&db[2]==0) {
    double e1[3],e2[3],n[3]; v_sub(b1,b0,e1); v_sub(b2,b0,e2); v_cross(e1,e2,n);
    if (n[0]==0&&n[1]==0&&n[2]==0) { v_sub(a1,a0,e1); v_sub(a2,a0,e2); v_cross(e1,e2,n); }
    if (n[0]==0&&n[1]==0&&n[2]==0) {   /* two zero-area triangles: segments in space, which meet only if some pair of edges touches */
      const double* av[3]={a0,a1,a2}; const double* bv[3]={b0,b1,b2};
      for (int i=0;i<3;i++) for (int j=0;j<3;j++) if (seg_seg_dist2(av[i],av[(i+1)%3],bv[j],bv[(j+1)%3])==0.0) return 1;
      return 0;
    }
    return tri_tri_coplanar(A,B,n);
  }
  /* edges of A against B */
  if (seg_tri(a0,a1,b0,b1,b2,da[0],da[1])) return 1;
  if (seg_tri(a1,a2,b0,b1,b2,da[1],da[2])) return 1;
  if (seg_tri(a2,a0,b0,b1,b2,da[2],da[0])) return 1;
  if (seg_tri(b0,b1,a0,a1,a2,db[0],db[1])) return 1;
  if (seg_tri(b1,b2,a0,a1,a2,db[1],db[2])) return 1;
  if (seg_tri(b2,b0,a0,a1,a2,db[2],db[0])) return 1;
  return 0;
}

/* closest point on triangle to p (region walk), returns squared distance */
static double point_tri_dist2(const double* p, const double* a, const double* b, const double* c) {
  double ab[3],ac[3],ap[3]; v_sub(b,a,ab); v_sub(c,a,ac); v_sub(p,a,ap);
  double d1=v_dot(ab,ap), d2=v_dot(ac,ap);
  if (d1<=0 && d2<=0) return v_dot(ap,ap);
  double bp[3]; v_sub(p,b,bp);
  double d3=v_dot(ab,bp), d4=v_dot(ac,bp);
  if (d3>=0 && d4<=d3) return v_dot(bp,bp);
  double vc=d1*d4-d3*d2;
  if (vc<=0 && d1>=0 && d3<=0) { double v=d1/(d1-d3); double q[3]; v_madd(a,ab,v,q); return v_dist2(p,q); }
  double cp[3]; v_sub(p,c,cp);
  double d5=v_dot(ab,cp), d6=v_dot(ac,cp);
  if (d6>=0 && d5<=d6) return v_dot(cp,cp);
  double vb=d5*d2-d1*d6;
  if (vb<=0 && d2>=0 && d6<=0) { double w=d2/(d2-d6); double q[3]; v_madd(a,ac,w,q); return v_dist2(p,q); }
  double va=d3*d6-d5*d4;
  if (va<=0 && (d4-d3)>=0 && (d5-d6)>=0) {
    double w=(d4-d3)/((d4-d3)+(d5-d6)); double bc[3],q[3]; v_sub(c,b,bc); v_madd(b,bc,w,q); return v_dist2(p,q); }
  /* interior: distance to plane */
  double n[3]; v_cross(ab,ac,n);
  double nn=v_dot(n,n);
  if (nn==0) { /* degenerate triangle: min over edges handled by callers via seg tests; fall back to vertices */
    double m=v_dot(ap,ap), t=v_dot(bp,bp); if (t<m) m=t; t=v_dot(cp,cp); if (t<m) m=t; return m; }
  double h=v_dot(ap,n);
  return h*h/nn;
}
double ko_point_tri_distance(const double p[3], const double t[9]) { return sqrt(point_tri_dist2(p,t,t+3,t+6)); }

/* closest points between two segments, squared distance */
static double seg_seg_dist2(const double* p1, const double* q1, const double* p2, const double* q2) {
  double d1[3],d2[3],r[3]; v_sub(q1,p1,d1); v_sub(q2,p2,d2); v_sub(p1,p2,r);
  double a=v_dot(d1,d1), e=v_dot(d2,d2), f=v_dot(d2,r);
  double s,t;
  if (a==0 && e==0) return v_dot(r,r);
  if (a==0) { s=0; t=f/e; if (t<0) t=0; if (t>1) t=1; }
  else {
    double c=v_dot(d1,r);
    if (e==0) { t=0; s=-c/a; if (s<0) s=0; if (s>1) s=1; }
    else {
      double b=v_dot(d1,d2), denom=a*e-b*b;
      if (denom>0) { s=(b*f-c*e)/denom; if (s<0) s=0; if (s>1) s=1; } else s=0;
      t=(b*s+f)/e;
      if (t<0) { t=0; s=-c/a; if (s<0) s=0; if (s>1) s=1; }
      else if (t>1) { t=1; s=(b-c)/a; if (s<0) s=0; if (s>1) s=1; }
    }
  }
  double c1[3],c2[3]; v_madd(p1,d1,s,c1); v_madd(p2,d2,t,c2);
  return v_dist2(c1,c2);
}
double ko_seg_seg_distance(const double p0[3], const double p1[3], const double q0[3], const double q1[3]) {
  return sqrt(seg_seg_dist2(p0,p1,q0,q1)); }

/* distance between two closed triangles: 0 if they intersect, else min over 9 edge pairs and 6 vertex-face pairs */
static double tri_tri_dist2(const double* A, const double* B) {
  if (ko_tri_tri_intersect(A,B)) return 0.0;
  double m=DBL_MAX, d;
  for (int i=0;i<3;i++) for (int j=0;j<3;j++) {
    d=seg_seg_dist2(A+3*i,A+3*((i+1)%3),B+3*j,B+3*((j+1)%3)); if (d<m) m=d; }
  for (int i=0;i<3;i++) { d=point_tri_dist2(A+3*i,B,B+3,B+6); if (d<m) m=d; }
  for (int i=0;i<3;i++) { d=point_tri_dist2(B+3*i,A,A+3,A+6); if (d<m) m=d; }
  return m;
}
double ko_tri_tri_distance(const double a[9], const double b[9]) { return sqrt(tri_tri_dist2(a,b)); }

/* ------------------------------------------------------------------ BVH build (median split, canonical) */
typedef struct { double c[3]; int idx; } cent_t;
static int g_sort_axis;
static int cent_cmp(const void* a, const void* b) {
  const cent_t* x=(const cent_t*)a; const cent_t* y=(const cent_t*)b;
  if (x->c[g_sort_axis]<y->c[g_sort_axis]) return -1;
  if (x->c[g_sort_axis]>y->c[g_sort_axis]) return 1;
  return (x->idx>y->idx)-(x->idx<y->idx);
}
typedef struct { node_t* nodes; int nnodes; cent_t* cents; const double* elo; const double* ehi; int leafsize; } build_t;

static int build_rec(build_t* b, int first, int count) {
  int me=b->nnodes++;
  node_t* nd=&b->nodes[me];
  for (int k=0;k<3;k++) { nd->lo[k]=DBL_MAX; nd->hi[k]=-DBL_MAX; }
  for (int i=first;i<first+count;i++) { int e=b->cents[i].idx;
    for (int k=0;k<3;k++) { if (b->elo[3*e+k]<nd->lo[k]) nd->lo[k]=b->elo[3*e+k]; if (b->ehi[3*e+k]>nd->hi[k]) nd->hi[k]=b->ehi[3*e+k]; } }
  nd->first=first; nd->count=count;
  if (count<=b->leafsize) { nd->left=-1; nd->right=-1; return me; }
  int ax=0; double ext=nd->hi[0]-nd->lo[0];
  for (int k=1;k<3;k++) if (nd->hi[k]-nd->lo[k]>ext) { ext=nd->hi[k]-nd->lo[k]; ax=k; }
  int half=count/2;
#ifdef KO_SAH
  /* CPU-baseline variant only (libkb_oracle_fast.so): binned surface-area heuristic, 16 bins x 3 axes over the centroid
   * bounds, the split a production BVH builder would choose.  Falls back to the median split when the centroids coincide.
   * The canonical tree of the checker (and of the roofline counts) stays the median split above. */
  if (count>2) {
    double clo[3]={DBL_MAX,DBL_MAX,DBL_MAX}, chi[3]={-DBL_MAX,-DBL_MAX,-DBL_MAX};
    for (int i=first;i<first+count;i++) for (int k=0;k<3;k++) { double c=b->cents[i].c[k]; if (c<clo[k]) clo[k]=c; if (c>chi[k]) chi[k]=c; }
    double bestc=DBL_MAX; int besta=-1, bestb=0;
    for (int a=0;a<3;a++) {
      double w=chi[a]-clo[a]; if (!(w>0)) continue;
      enum { NB=16 };
      int cnt[NB]; double blo[NB][3], bhi[NB][3];
      for (int i=0;i<NB;i++) { cnt[i]=0; for (int k=0;k<3;k++) { blo[i][k]=DBL_MAX; bhi[i][k]=-DBL_MAX; } }
      for (int i=first;i<first+count;i++) { int e=b->cents[i].idx; int bi=(int)((b->cents[i].c[a]-clo[a])/w*NB); if (bi>=NB) bi=NB-1;
        cnt[bi]++; for (int k=0;k<3;k++) { if (b->elo[3*e+k]<blo[bi][k]) blo[bi][k]=b->elo[3*e+k]; if (b->ehi[3*e+k]>bhi[bi][k]) bhi[bi][k]=b->ehi[3*e+k]; } }
      double ra[NB]; int rn[NB]; { double lo[3]={DBL_MAX,DBL_MAX,DBL_MAX}, hi[3]={-DBL_MAX,-DBL_MAX,-DBL_MAX}; int n=0;
        for (int i=NB-1;i>0;i--) { n+=cnt[i]; for (int k=0;k<3;k++) { if (blo[i][k]<lo[k]) lo[k]=blo[i][k]; if (bhi[i][k]>hi[k]) hi[k]=bhi[i][k]; }
          double d0=hi[0]-lo[0], d1=hi[1]-lo[1], d2=hi[2]-lo[2]; ra[i]= n? (d0*d1+d1*d2+d2*d0):0; rn[i]=n; } }
      double lo[3]={DBL_MAX,DBL_MAX,DBL_MAX}, hi[3]={-DBL_MAX,-DBL_MAX,-DBL_MAX}; int n=0;
      for (int i=0;i<NB-1;i++) { n+=cnt[i]; for (int k=0;k<3;k++) { if (blo[i][k]<lo[k]) lo[k]=blo[i][k]; if (bhi[i][k]>hi[k]) hi[k]=bhi[i][k]; }
        if (!n || !rn[i+1]) continue;
        double d0=hi[0]-lo[0], d1=hi[1]-lo[1], d2=hi[2]-lo[2];
        double c=(d0*d1+d1*d2+d2*d0)*n+ra[i+1]*rn[i+1];
        if (c<bestc) { bestc=c; besta=a; bestb=i; } }
    }
    if (besta>=0) {
      double w=chi[besta]-clo[besta]; int i=first, j=first+count-1;
      while (i<=j) { int bi=(int)((b->cents[i].c[besta]-clo[besta])/w*16); if (bi>=16) bi=15;
        if (bi<=bestb) i++; else { cent_t t=b->cents[i]; b->cents[i]=b->cents[j]; b->cents[j]=t; j--; } }
      half=i-first;
      if (half<=0 || half>=count) { besta=-1; half=count/2; }
    }
    if (besta<0) { g_sort_axis=ax; qsort(b->cents+first,count,sizeof(cent_t),cent_cmp); }
  } else
#endif
  {
  g_sort_axis=ax;
  qsort(b->cents+first,count,sizeof(cent_t),cent_cmp);
  }
  int l=build_rec(b,first,half);
  int r=build_rec(b,first+half,count-half);
  b->nodes[me].left=l; b->nodes[me].right=r;
  return me;
}

/* A triangle whose area is below 1e-12 of its longest edge squared cannot be told from a segment in fp64: the signs of the
 * orientation tests against its "plane" are rounding noise.  It is replaced by the segment between its two farthest vertices,
 * written as the exactly degenerate triangle (p, q, q), which the predicates handle as a segment.  The engine applies the same
 * rule when a mesh is added (kb_add_trimesh), so both sides see the same geometry. */
static void snap_sliver(double* t) {
  double e0[3],e1[3],e2[3],n[3]; v_sub(t+3,t,e0); v_sub(t+6,t,e1); v_sub(t+6,t+3,e2); v_cross(e0,e1,n);
  double l0=v_dot(e0,e0), l1=v_dot(e1,e1), l2=v_dot(e2,e2), L=fmax(l0,fmax(l1,l2));
  if (v_dot(n,n) > 1e-24*L*L) return;
  double p[3],q[3];
  if (L==l0) { memcpy(p,t,24); memcpy(q,t+3,24); } else if (L==l1) { memcpy(p,t,24); memcpy(q,t+6,24); } else { memcpy(p,t+3,24); memcpy(q,t+6,24); }
  memcpy(t,p,24); memcpy(t+3,q,24); memcpy(t+6,q,24);
}
static void geom_build_bvh(geom_t* g) {
  int n = (g->kind==G_MESH) ? g->nt : g->np;
  if (n<=0) { g->nodes=NULL; g->nnodes=0; return; }
  double* elo=(double*)malloc(sizeof(double)*3*n); double* ehi=(double*)malloc(sizeof(double)*3*n);
  cent_t* cents=(cent_t*)malloc(sizeof(cent_t)*n);
  for (int e=0;e<n;e++) {
    cents[e].idx=e;
    if (g->kind==G_MESH) {
      for (int k=0;k<3;k++) {
        double a=g->verts[3*g->tris[3*e]+k], b=g->verts[3*g->tris[3*e+1]+k], c=g->verts[3*g->tris[3*e+2]+k];
        elo[3*e+k]=fmin(a,fmin(b,c)); ehi[3*e+k]=fmax(a,fmax(b,c));
        cents[e].c[k]=(a+b+c)/3.0; }
    } else {
      for (int k=0;k<3;k++) { elo[3*e+k]=g->pts[3*e+k]-g->rad[e]; ehi[3*e+k]=g->pts[3*e+k]+g->rad[e]; cents[e].c[k]=g->pts[3*e+k]; }
    }
  }
  build_t b; b.leafsize=(g->kind==G_MESH)?1:CLOUD_LEAF;
  b.nodes=(node_t*)malloc(sizeof(node_t)*(2*(size_t)n)); b.nnodes=0; b.cents=cents; b.elo=elo; b.ehi=ehi;
  build_rec(&b,0,n);
  g->nodes=b.nodes; g->nnodes=b.nnodes;
  g->perm=(int*)malloc(sizeof(int)*n);
  for (int i=0;i<n;i++) g->perm[i]=cents[i].idx;
  if (g->kind==G_MESH) {
    g->tv=(double*)malloc(sizeof(double)*9*(size_t)n);
    for (int i=0;i<n;i++) { int e=g->perm[i];
      for (int v=0;v<3;v++) for (int k=0;k<3;k++) g->tv[9*(size_t)i+3*v+k]=g->verts[3*g->tris[3*e+v]+k];
      snap_sliver(g->tv+9*(size_t)i); }
  } else {
    double* p2=(double*)malloc(sizeof(double)*3*(size_t)n); double* r2=(double*)malloc(sizeof(double)*n);
    for (int i=0;i<n;i++) { int e=g->perm[i]; memcpy(p2+3*(size_t)i,g->pts+3*(size_t)e,3*sizeof(double)); r2[i]=g->rad[e]; }
    free(g->pts); free(g->rad); g->pts=p2; g->rad=r2;
  }
  memcpy(g->lo,g->nodes[0].lo,sizeof(g->lo)); memcpy(g->hi,g->nodes[0].hi,sizeof(g->hi));
  g->rmax=0; if (g->kind!=G_MESH) for (int i=0;i<n;i++) if (g->rad[i]>g->rmax) g->rmax=g->rad[i];
  free(elo); free(ehi); free(cents);
}
static inline int node_right(const geom_t* g, int i) { return g->nodes[i].right; }

/* ------------------------------------------------------------------ world construction */
ko_world* ko_create(void) { ko_world* w=(ko_world*)calloc(1,sizeof(ko_world)); w->selfcol_default=1; return w; }

static void geom_free(geom_t* g) { free(g->verts); free(g->tris); free(g->tv); free(g->pts); free(g->rad); free(g->perm); free(g->nodes); }
void ko_destroy(ko_world* w) {
  if (!w) return;
  for (int i=0;i<w->ngeoms;i++) geom_free(&w->geoms[i]);
  free(w->geoms); free(w->terrains); free(w->objects); free(w->objT);
  free(w->parents); free(w->linktype); free(w->axis); free(w->T0); free(w->qmin); free(w->qmax); free(w->linkgeom);
  free(w->jtype); free(w->jlink); free(w->jbase);
  for (int i=0;i<w->ndrv;i++) { free(w->drivers[i].links); free(w->drivers[i].scale); free(w->drivers[i].offset); }
  free(w->drivers); free(w->selfcol); free(w->mask);
  free(w);
}
static geom_t* new_geom(ko_world* w) {
  if (w->ngeoms==w->capgeoms) { w->capgeoms=w->capgeoms? 2*w->capgeoms:16; w->geoms=(geom_t*)realloc(w->geoms,sizeof(geom_t)*w->capgeoms); }
  geom_t* g=&w->geoms[w->ngeoms++]; memset(g,0,sizeof(*g)); return g;
}
int ko_add_trimesh(ko_world* w, const double* verts, int nv, const int32_t* tris, int nt, double margin) {
  geom_t* g=new_geom(w); g->kind=(nt>0)?G_MESH:G_EMPTY; g->margin=margin; g->nv=nv; g->nt=nt;
  g->verts=(double*)malloc(sizeof(double)*3*(nv>0?nv:1)); memcpy(g->verts,verts,sizeof(double)*3*nv);
  g->tris=(int32_t*)malloc(sizeof(int32_t)*3*(nt>0?nt:1)); memcpy(g->tris,tris,sizeof(int32_t)*3*nt);
  geom_build_bvh(g);
  return w->ngeoms-1;
}
int ko_add_pointcloud(ko_world* w, const double* pts, int n, const double* radius, double margin) {
  geom_t* g=new_geom(w); g->kind=(n>0)?G_CLOUD:G_EMPTY; g->margin=margin; g->np=n;
  g->pts=(double*)malloc(sizeof(double)*3*(n>0?n:1)); memcpy(g->pts,pts,sizeof(double)*3*n);
  g->rad=(double*)calloc(n>0?n:1,sizeof(double)); if (radius) memcpy(g->rad,radius,sizeof(double)*n);
  geom_build_bvh(g);
  return w->ngeoms-1;
}
int ko_add_primitive(ko_world* w, int type, const double* params, double margin) {
  if (type==KO_PRIM_TRIANGLE) { int32_t idx[3]={0,1,2}; return ko_add_trimesh(w,params,3,idx,1,margin); }   /* one-triangle mesh */
  if (type==KO_PRIM_SEGMENT) {          /* Segment3D = the zero-area triangle (a,b,b); ko_tri_tri_* treat it as the segment */
    if (params[0]==params[3] && params[1]==params[4] && params[2]==params[5]) return -1;
    int32_t idx[3]={0,1,1}; return ko_add_trimesh(w,params,2,idx,1,margin); }
  if (type==KO_PRIM_BOX || type==KO_PRIM_AABB) {
    /* solid box (GeometricPrimitive3D Box3D / AABB3D; common primitives of Cpp/docs/Manual-Geometry.md:241-250): its surface
     * as 12 triangles + the solid descriptor.  BOX params: centre(3), axes as the columns of a row-major 3x3 (9), half dims(3);
     * AABB params: lo(3), hi(3). */
    double c[3], R[9]={1,0,0,0,1,0,0,0,1}, h[3];
    if (type==KO_PRIM_AABB) { for (int k=0;k<3;k++) { c[k]=0.5*(params[k]+params[3+k]); h[k]=0.5*(params[3+k]-params[k]); } }
    else { memcpy(c,params,24); memcpy(R,params+3,72); memcpy(h,params+12,24); }
    for (int k=0;k<3;k++) if (!(h[k]>=0)) return -1;
    int nzero=(h[0]==0)+(h[1]==0)+(h[2]==0);
    if (nzero>=2) return -1;            /* a segment or a point, not a box */
    double v[24]; int nvv=0;
    for (int sz=-1;sz<=1;sz+=2) for (int sy=-1;sy<=1;sy+=2) for (int sx=-1;sx<=1;sx+=2) {
      double l[3]={sx*h[0],sy*h[1],sz*h[2]};
      for (int k=0;k<3;k++) v[3*nvv+k]=R[3*k]*l[0]+R[3*k+1]*l[1]+R[3*k+2]*l[2]+c[k];
      nvv++; }
    /* vertex index = (x>0) + 2 (y>0) + 4 (z>0) */
    static const int32_t T[36]={0,2,3, 0,3,1,  4,5,7, 4,7,6,  0,1,5, 0,5,4,  2,6,7, 2,7,3,  0,4,6, 0,6,2,  1,3,7, 1,7,5};
    /* faces of T in the order z-, z+, y-, y+, x-, x+; a flat box keeps only the two faces across its zero dimension (the other
     * four have no area) */
    int32_t TT[36]; int ntt=0;
    for (int f=0;f<6;f++) { int axis=2-f/2; if (nzero==1 && h[axis]!=0) continue; for (int k=6*f;k<6*f+6;k++) TT[ntt++]=T[k]; }
    int gi=ko_add_trimesh(w,v,8,TT,ntt/3,margin);
    geom_t* g=&w->geoms[gi]; g->solid=1; memcpy(g->bc,c,24); memcpy(g->bR,R,72); memcpy(g->bh,h,24);
    return gi;
  }
  double r = (type==KO_PRIM_SPHERE) ? params[3] : 0.0;
  if (type!=KO_PRIM_POINT && type!=KO_PRIM_SPHERE) return -1;
  int gi=ko_add_pointcloud(w,params,1,&r,margin);
  w->geoms[gi].kind=G_PRIM;
  return gi;
}
int ko_add_terrain(ko_world* w, int geom) {
  w->terrains=(int*)realloc(w->terrains,sizeof(int)*(w->nterr+1)); w->terrains[w->nterr]=geom; return w->nterr++; }
int ko_add_rigid_object(ko_world* w, int geom, const double T[12]) {
  w->objects=(int*)realloc(w->objects,sizeof(int)*(w->nobj+1)); w->objT=(xf_t*)realloc(w->objT,sizeof(xf_t)*(w->nobj+1));
  w->objects[w->nobj]=geom; xf_from12(T,&w->objT[w->nobj]); return w->nobj++; }
int ko_robot_create(ko_world* w, int L, const int32_t* parents, const uint8_t* linktype,
                    const double* axis, const double* T0, const double* qmin, const double* qmax) {
  if (w->L) return -1;
  w->L=L;
  w->parents=(int32_t*)malloc(sizeof(int32_t)*L); memcpy(w->parents,parents,sizeof(int32_t)*L);
  w->linktype=(uint8_t*)malloc(L); memcpy(w->linktype,linktype,L);
  w->axis=(double*)malloc(sizeof(double)*3*L); memcpy(w->axis,axis,sizeof(double)*3*L);
  w->T0=(xf_t*)malloc(sizeof(xf_t)*L); for (int i=0;i<L;i++) xf_from12(T0+12*i,&w->T0[i]);
  w->qmin=(double*)malloc(sizeof(double)*L); memcpy(w->qmin,qmin,sizeof(double)*L);
  w->qmax=(double*)malloc(sizeof(double)*L); memcpy(w->qmax,qmax,sizeof(double)*L);
  w->linkgeom=(int*)malloc(sizeof(int)*L); for (int i=0;i<L;i++) w->linkgeom[i]=-1;
  /* default joints: one Normal joint per link (Robot.cpp default when no "joint" lines) */
  w->nj=L; w->jtype=(uint8_t*)malloc(L); w->jlink=(int32_t*)malloc(sizeof(int32_t)*L); w->jbase=(int32_t*)malloc(sizeof(int32_t)*L);
  for (int i=0;i<L;i++) { w->jtype[i]=KO_JOINT_NORMAL; w->jlink[i]=i; w->jbase[i]=parents[i]; }
  w->selfcol=(uint8_t*)calloc((size_t)L*L,1);
  for (int i=0;i<L;i++) if (parents[i]>=i) return -2;
  return 0;
}
int ko_robot_set_link_geometry(ko_world* w, int link, int geom) { if (link<0||link>=w->L) return -1; w->linkgeom[link]=geom; return 0; }
/* links a joint drives: the chain from its base link (exclusive) down to its link (inclusive), in root-to-tip order
 * (RobotModel::GetJointIndices, Cpp/Modeling/Robot.cpp:2120-2144).  Returns the count, or -1 if the chain never meets the base. */
static int joint_indices(const ko_world* w, int j, int idx[6]) {
  int t=w->jtype[j], n=0, tmp[8], link=w->jlink[j];
  if (t==KO_JOINT_WELD||t==KO_JOINT_NORMAL||t==KO_JOINT_SPIN) { idx[0]=link; return 1; }
  while (link!=w->jbase[j]) { if (link<0||n>=6) return -1; tmp[n++]=link; link=w->parents[link]; }
  for (int i=0;i<n;i++) idx[i]=tmp[n-1-i];
  return n;
}
int ko_robot_set_joints(ko_world* w, int nj, const uint8_t* jtype, const int32_t* jlink, const int32_t* jbase) {
  free(w->jtype); free(w->jlink); free(w->jbase); w->nj=nj;
  w->jtype=(uint8_t*)malloc(nj>0?nj:1); memcpy(w->jtype,jtype,nj);
  w->jlink=(int32_t*)malloc(sizeof(int32_t)*(nj>0?nj:1)); memcpy(w->jlink,jlink,sizeof(int32_t)*nj);
  w->jbase=(int32_t*)malloc(sizeof(int32_t)*(nj>0?nj:1));
  for (int i=0;i<nj;i++) w->jbase[i]=jbase?jbase[i]:w->parents[jlink[i]];
  /* the reference asserts the link layout of multi-link joints (Interpolate.cpp:24-26,231-236): translations first, then
   * rotations about z, y, x */
  for (int i=0;i<nj;i++) {
    int idx[6], n=joint_indices(w,i,idx), t=jtype[i];
    if (t==KO_JOINT_FLOATING) { if (n!=6) return -1;
      for (int k=0;k<3;k++) if (w->linktype[idx[k]]!=KO_PRISMATIC || w->linktype[idx[3+k]]!=KO_REVOLUTE) return -1;
      if (w->axis[3*idx[3]+2]!=1.0 || w->axis[3*idx[4]+1]!=1.0 || w->axis[3*idx[5]]!=1.0) return -1; }
    else if (t==KO_JOINT_BALLANDSOCKET) { if (n!=3) return -1;
      for (int k=0;k<3;k++) if (w->linktype[idx[k]]!=KO_REVOLUTE) return -1;
      if (w->axis[3*idx[0]+2]!=1.0 || w->axis[3*idx[1]+1]!=1.0 || w->axis[3*idx[2]]!=1.0) return -1; }
    else if (t==KO_JOINT_FLOATINGPLANAR) { if (n!=3 || w->linktype[idx[2]]!=KO_REVOLUTE) return -1; }
    else if (t==KO_JOINT_CLOSED) return -1;
  }
  return 0; }
int ko_robot_add_affine_driver(ko_world* w, int n, const int32_t* links, const double* scale, const double* offset, double dmin, double dmax) {
  w->drivers=(driver_t*)realloc(w->drivers,sizeof(driver_t)*(w->ndrv+1));
  driver_t* d=&w->drivers[w->ndrv++]; d->n=n; d->dmin=dmin; d->dmax=dmax;
  d->links=(int32_t*)malloc(sizeof(int32_t)*n); memcpy(d->links,links,sizeof(int32_t)*n);
  d->scale=(double*)malloc(sizeof(double)*n); d->offset=(double*)malloc(sizeof(double)*n);
  for (int i=0;i<n;i++) { d->scale[i]=scale?scale[i]:1.0; d->offset[i]=offset?offset[i]:0.0; }
  return w->ndrv-1;
}
static int geom_empty(const ko_world* w, int g) { return g<0 || w->geoms[g].kind==G_EMPTY; }

/* a7: default self-collision set = all i<j, both non-empty, neither the other's parent */
static void init_all_self_collisions(ko_world* w) {
  int L=w->L;
  for (int i=0;i<L;i++) for (int j=i+1;j<L;j++) {
    int en = !geom_empty(w,w->linkgeom[i]) && !geom_empty(w,w->linkgeom[j]) && w->parents[j]!=i && w->parents[i]!=j;
    w->selfcol[i*L+j]=(uint8_t)en; }
}
int ko_robot_set_self_collision(ko_world* w, int i, int j, int enabled) {
  if (i>j) { int t=i; i=j; j=t; }
  if (i==j || i<0 || j>=w->L) return -1;
  if (w->selfcol_default) { init_all_self_collisions(w); w->selfcol_default=0; }
  if (enabled && (geom_empty(w,w->linkgeom[i])||geom_empty(w,w->linkgeom[j]))) enabled=0;
  w->selfcol[i*w->L+j]=(uint8_t)(enabled!=0); return 0;
}
int ko_num_ids(const ko_world* w) { return w->nterr+w->nobj+(w->L?1+w->L:0); }
int ko_set_pair_mask(ko_world* w, const uint8_t* mask, int n_ids) {
  if (n_ids!=ko_num_ids(w)) return -1;
  free(w->mask); w->mask=(uint8_t*)malloc((size_t)n_ids*n_ids); memcpy(w->mask,mask,(size_t)n_ids*n_ids);
  w->nids=n_ids; w->mask_user=1; return 0;
}
/* a6: WorldPlannerSettings::InitializeDefault, PlannerSettings.cpp:16-41 */
static void init_default_mask(ko_world* w) {
  int n=ko_num_ids(w); w->nids=n;
  w->mask=(uint8_t*)malloc((size_t)n*n); memset(w->mask,1,(size_t)n*n);
  for (int i=0;i<n;i++) w->mask[i*n+i]=0;
  if (w->L) {
    int k=w->nterr+w->nobj, base=k+1, L=w->L;
    w->mask[k*n+k]=1;
    for (int j=0;j<L;j++) { w->mask[(base+j)*n+k]=0; w->mask[k*n+base+j]=0; }
    for (int j=0;j<L;j++) for (int m=0;m<L;m++)
      w->mask[(base+j)*n+base+m] = (j<m) ? w->selfcol[j*L+m] : 0;   /* selfCollisions(j,k) upper triangular */
    for (int j=0;j<L;j++) if (w->parents[j]==-1)
      for (int t=0;t<w->nterr;t++) { w->mask[(base+j)*n+t]=0; w->mask[t*n+base+j]=0; }
  }
}
int ko_finalize(ko_world* w) {
  if (w->L && w->selfcol_default) { init_all_self_collisions(w); w->selfcol_default=0; }
  if (!w->mask_user) { free(w->mask); init_default_mask(w); }
  w->finalized=1; return 0;
}
int ko_get_pair_mask(const ko_world* w, uint8_t* out) { memcpy(out,w->mask,(size_t)w->nids*w->nids); return w->nids; }

/* ------------------------------------------------------------------ FK, limits */
static void axis_angle(const double* w, double th, double* R) {
  double c=cos(th), s=sin(th), v=1.0-c;
  R[0]=c+v*w[0]*w[0];      R[1]=v*w[0]*w[1]-s*w[2]; R[2]=v*w[0]*w[2]+s*w[1];
  R[3]=v*w[1]*w[0]+s*w[2]; R[4]=c+v*w[1]*w[1];      R[5]=v*w[1]*w[2]-s*w[0];
  R[6]=v*w[2]*w[0]-s*w[1]; R[7]=v*w[2]*w[1]+s*w[0]; R[8]=c+v*w[2]*w[2];
}
static void fk_links(const ko_world* w, const double* q, xf_t* T) {
  for (int i=0;i<w->L;i++) {
    xf_t loc; xf_identity(&loc);
    if (w->linktype[i]==KO_PRISMATIC) { loc.t[0]=q[i]*w->axis[3*i]; loc.t[1]=q[i]*w->axis[3*i+1]; loc.t[2]=q[i]*w->axis[3*i+2]; }
    else axis_angle(w->axis+3*i,q[i],loc.R);
    xf_t rel; xf_mul(&w->T0[i],&loc,&rel);
    if (w->parents[i]<0) T[i]=rel; else xf_mul(&T[w->parents[i]],&rel,&T[i]);
  }
}
void ko_fk(const ko_world* w, const double* q, double* T_out) {
  xf_t* T=(xf_t*)malloc(sizeof(xf_t)*w->L); fk_links(w,q,T);
  for (int i=0;i<w->L;i++) xf_to12(&T[i],T_out+12*i);
  free(T);
}
int ko_check_joint_limits(const ko_world* w, const double* q) {
  for (int i=0;i<w->nj;i++) if (w->jtype[i]==KO_JOINT_NORMAL || w->jtype[i]==KO_JOINT_WELD) {
    int k=w->jlink[i]; if (q[k]<w->qmin[k] || q[k]>w->qmax[k]) return 0; }
  for (int i=0;i<w->ndrv;i++) { const driver_t* d=&w->drivers[i];
    double v=0; for (int j=0;j<d->n;j++) v+=(q[d->links[j]]-d->offset[j])/d->scale[j];
    v/=d->n; if (v<d->dmin || v>d->dmax) return 0; }
  return 1;
}

/* ------------------------------------------------------------------ BV tests and traversals */
/* world AABB of a transformed local box, inflated by margin: AnyCollisionGeometry3D::GetAABB (O(1), loose) */
static void box_world_aabb(const double* lo, const double* hi, const xf_t* T, double m, double* bmin, double* bmax) {
  double c[3]={0.5*(lo[0]+hi[0]),0.5*(lo[1]+hi[1]),0.5*(lo[2]+hi[2])}, h[3]={0.5*(hi[0]-lo[0]),0.5*(hi[1]-lo[1]),0.5*(hi[2]-lo[2])};
  double wc[3]; xf_apply(T,c,wc);
  for (int i=0;i<3;i++) { double e=fabs(T->R[3*i])*h[0]+fabs(T->R[3*i+1])*h[1]+fabs(T->R[3*i+2])*h[2]+m; bmin[i]=wc[i]-e; bmax[i]=wc[i]+e; }
}
void ko_geom_aabb(const ko_world* w, int g, const double T12[12], double bmin[3], double bmax[3]) {
  xf_t T; xf_from12(T12,&T); const geom_t* G=&w->geoms[g];
  box_world_aabb(G->lo,G->hi,&T,G->margin,bmin,bmax);
}
static inline int aabb_overlap(const double* alo, const double* ahi, const double* blo, const double* bhi) {
  return alo[0]<=bhi[0] && blo[0]<=ahi[0] && alo[1]<=bhi[1] && blo[1]<=ahi[1] && alo[2]<=bhi[2] && blo[2]<=ahi[2]; }

/* OBB-OBB separating-axis test; box B is expressed in A's frame by Tab; A half-extents inflated by tol. */
static int obb_overlap(const node_t* a, const node_t* b, const xf_t* Tab, double tol) {
  double ca[3],ha[3],cb[3],hb[3];
  for (int i=0;i<3;i++) { ca[i]=0.5*(a->lo[i]+a->hi[i]); ha[i]=0.5*(a->hi[i]-a->lo[i])+tol; cb[i]=0.5*(b->lo[i]+b->hi[i]); hb[i]=0.5*(b->hi[i]-b->lo[i]); }
  double cbA[3]; xf_apply(Tab,cb,cbA);
  double t[3]; v_sub(cbA,ca,t);
  const double* R=Tab->R; double AR[9];
  for (int i=0;i<9;i++) AR[i]=fabs(R[i])+1e-12;
  for (int i=0;i<3;i++) if (fabs(t[i]) > ha[i]+hb[0]*AR[3*i]+hb[1]*AR[3*i+1]+hb[2]*AR[3*i+2]) return 0;
  for (int j=0;j<3;j++) if (fabs(t[0]*R[j]+t[1]*R[3+j]+t[2]*R[6+j]) > hb[j]+ha[0]*AR[j]+ha[1]*AR[3+j]+ha[2]*AR[6+j]) return 0;
  for (int i=0;i<3;i++) { int i1=(i+1)%3, i2=(i+2)%3;
    for (int j=0;j<3;j++) { int j1=(j+1)%3, j2=(j+2)%3;
      double ra=ha[i1]*AR[3*i2+j]+ha[i2]*AR[3*i1+j];
      double rb=hb[j1]*AR[3*i+j2]+hb[j2]*AR[3*i+j1];
      if (fabs(t[i2]*R[3*i1+j]-t[i1]*R[3*i2+j]) > ra+rb) return 0; } }
  return 1;
}
/* lower bound on the distance between two boxes (B in A's frame): max over the 6 face axes of the gap */
static double obb_dist_lb(const node_t* a, const node_t* b, const xf_t* Tab) {
  double ca[3],ha[3],cb[3],hb[3];
  for (int i=0;i<3;i++) { ca[i]=0.5*(a->lo[i]+a->hi[i]); ha[i]=0.5*(a->hi[i]-a->lo[i]); cb[i]=0.5*(b->lo[i]+b->hi[i]); hb[i]=0.5*(b->hi[i]-b->lo[i]); }
  double cbA[3]; xf_apply(Tab,cb,cbA); double t[3]; v_sub(cbA,ca,t);
  const double* R=Tab->R; double g2=0, best=0;
  /* A's axes: per-axis gaps combine in quadrature (AABB-AABB distance with B's AABB in A's frame) */
  for (int i=0;i<3;i++) { double e=hb[0]*fabs(R[3*i])+hb[1]*fabs(R[3*i+1])+hb[2]*fabs(R[3*i+2]);
    double g=fabs(t[i])-ha[i]-e; if (g>0) g2+=g*g; }
  best=sqrt(g2); g2=0;
  for (int j=0;j<3;j++) { double e=ha[0]*fabs(R[j])+ha[1]*fabs(R[3+j])+ha[2]*fabs(R[6+j]);
    double g=fabs(t[0]*R[j]+t[1]*R[3+j]+t[2]*R[6+j])-hb[j]-e; if (g>0) g2+=g*g; }
  double b2=sqrt(g2); return b2>best?b2:best;
}
/* Lower bound on the signed element distance below two nodes.  Sphere radii are inside the node boxes, so a positive
 * box gap bounds the ball distance; when the boxes touch the balls may interpenetrate by at most the radii. */
static inline double signed_lb(double box_lb, double rsum) { return box_lb>0 ? box_lb : -rsum; }
static inline double node_size2(const node_t* n) { double d[3]; v_sub(n->hi,n->lo,d); return v_dot(d,d); }

typedef struct { const geom_t* A; const geom_t* B; xf_t Ta, Tb, Tab; double tol; double rsum; ko_counts* cnt; } pairq_t;

/* leaf-vs-leaf: returns 1 if any element pair is within tol (tol==0: intersects) */
static int leaf_collide(pairq_t* q, const node_t* a, const node_t* b) {
  const geom_t *A=q->A, *B=q->B;
  if (A->kind==G_MESH && B->kind==G_MESH) {
    double ta[9], tb[9];
    for (int v=0;v<3;v++) { xf_apply(&q->Ta,A->tv+9*(size_t)a->first+3*v,ta+3*v); xf_apply(&q->Tb,B->tv+9*(size_t)b->first+3*v,tb+3*v); }
    if (q->cnt) q->cnt->n_tri++;
    if (q->tol==0) return ko_tri_tri_intersect(ta,tb);
    return tri_tri_dist2(ta,tb) <= q->tol*q->tol;
  }
  if (A->kind==G_MESH) { /* triangle vs point-spheres */
    double ta[9]; for (int v=0;v<3;v++) xf_apply(&q->Ta,A->tv+9*(size_t)a->first+3*v,ta+3*v);
    for (int i=b->first;i<b->first+b->count;i++) { double p[3]; xf_apply(&q->Tb,B->pts+3*(size_t)i,p);
      if (q->cnt) q->cnt->n_pt++;
      double r=B->rad[i]+q->tol; if (point_tri_dist2(p,ta,ta+3,ta+6) <= r*r) return 1; }
    return 0;
  }
  if (B->kind==G_MESH) {
    double tb[9]; for (int v=0;v<3;v++) xf_apply(&q->Tb,B->tv+9*(size_t)b->first+3*v,tb+3*v);
    for (int i=a->first;i<a->first+a->count;i++) { double p[3]; xf_apply(&q->Ta,A->pts+3*(size_t)i,p);
      if (q->cnt) q->cnt->n_pt++;
      double r=A->rad[i]+q->tol; if (point_tri_dist2(p,tb,tb+3,tb+6) <= r*r) return 1; }
    return 0;
  }
  for (int i=a->first;i<a->first+a->count;i++) { double p[3]; xf_apply(&q->Ta,A->pts+3*(size_t)i,p);
    for (int j=b->first;j<b->first+b->count;j++) { double s[3]; xf_apply(&q->Tb,B->pts+3*(size_t)j,s);
      if (q->cnt) q->cnt->n_pt++;
      double r=A->rad[i]+B->rad[j]+q->tol; if (v_dist2(p,s) <= r*r) return 1; } }
  return 0;
}
/* leaf-vs-leaf signed-for-spheres distance (geometric distance minus radii) */
static double leaf_distance(pairq_t* q, const node_t* a, const node_t* b) {
  const geom_t *A=q->A, *B=q->B; double best=DBL_MAX;
  if (A->kind==G_MESH && B->kind==G_MESH) {
    double ta[9], tb[9];
    for (int v=0;v<3;v++) { xf_apply(&q->Ta,A->tv+9*(size_t)a->first+3*v,ta+3*v); xf_apply(&q->Tb,B->tv+9*(size_t)b->first+3*v,tb+3*v); }
    if (q->cnt) q->cnt->n_tri++;
    return sqrt(tri_tri_dist2(ta,tb));
  }
  if (A->kind==G_MESH || B->kind==G_MESH) {
    const geom_t* M = A->kind==G_MESH?A:B; const geom_t* C = A->kind==G_MESH?B:A;
    const node_t* mn = A->kind==G_MESH?a:b; const node_t* cn = A->kind==G_MESH?b:a;
    const xf_t* Tm = A->kind==G_MESH?&q->Ta:&q->Tb; const xf_t* Tc = A->kind==G_MESH?&q->Tb:&q->Ta;
    double t[9]; for (int v=0;v<3;v++) xf_apply(Tm,M->tv+9*(size_t)mn->first+3*v,t+3*v);
    for (int i=cn->first;i<cn->first+cn->count;i++) { double p[3]; xf_apply(Tc,C->pts+3*(size_t)i,p);
      if (q->cnt) q->cnt->n_pt++;
      double d=sqrt(point_tri_dist2(p,t,t+3,t+6))-C->rad[i]; if (d<best) best=d; }
    return best;
  }
  for (int i=a->first;i<a->first+a->count;i++) { double p[3]; xf_apply(&q->Ta,A->pts+3*(size_t)i,p);
    for (int j=b->first;j<b->first+b->count;j++) { double s[3]; xf_apply(&q->Tb,B->pts+3*(size_t)j,s);
      if (q->cnt) q->cnt->n_pt++;
      double d=sqrt(v_dist2(p,s))-A->rad[i]-B->rad[j]; if (d<best) best=d; } }
  return best;
}

/* canonical boolean traversal: simultaneous descent, descend-larger-first, first-contact exit */
static int collide_rec(pairq_t* q, int ia, int ib) {
  const node_t* a=&q->A->nodes[ia]; const node_t* b=&q->B->nodes[ib];
  if (q->cnt) q->cnt->n_node++;
  if (!obb_overlap(a,b,&q->Tab,q->tol)) return 0;
  int la=a->left<0, lb=b->left<0;
  if (la && lb) return leaf_collide(q,a,b);
  if (lb || (!la && node_size2(a)>=node_size2(b))) {
    int l=a->left; if (collide_rec(q,l,ib)) return 1;
    return collide_rec(q,node_right(q->A,ia),ib);
  } else {
    int l=b->left; if (collide_rec(q,ia,l)) return 1;
    return collide_rec(q,ia,node_right(q->B,ib));
  }
}
/* branch and bound distance; *best is the running minimum (starts at the upper bound) */
static void distance_rec(pairq_t* q, int ia, int ib, double* best) {
  const node_t* a=&q->A->nodes[ia]; const node_t* b=&q->B->nodes[ib];
  if (q->cnt) q->cnt->n_node++;
  /* elements may be spheres: their radius is already inside the node boxes */
  if (signed_lb(obb_dist_lb(a,b,&q->Tab),q->rsum) >= *best) return;
  int la=a->left<0, lb=b->left<0;
  if (la && lb) { double d=leaf_distance(q,a,b); if (d<*best) *best=d; return; }
  if (lb || (!la && node_size2(a)>=node_size2(b))) {
    int l=a->left, r=node_right(q->A,ia);
    double dl=obb_dist_lb(&q->A->nodes[l],b,&q->Tab), dr=obb_dist_lb(&q->A->nodes[r],b,&q->Tab);
    if (dl<=dr) { distance_rec(q,l,ib,best); distance_rec(q,r,ib,best); } else { distance_rec(q,r,ib,best); distance_rec(q,l,ib,best); }
  } else {
    int l=b->left, r=node_right(q->B,ib);
    double dl=obb_dist_lb(a,&q->B->nodes[l],&q->Tab), dr=obb_dist_lb(a,&q->B->nodes[r],&q->Tab);
    if (dl<=dr) { distance_rec(q,ia,l,best); distance_rec(q,ia,r,best); } else { distance_rec(q,ia,r,best); distance_rec(q,ia,l,best); }
  }
}
static void pairq_init(pairq_t* q, const geom_t* A, const xf_t* Ta, const geom_t* B, const xf_t* Tb, double tol, ko_counts* cnt) {
  q->A=A; q->B=B; q->Ta=*Ta; q->Tb=*Tb; xf_mul_inv_a(Ta,Tb,&q->Tab); q->tol=tol; q->rsum=A->rmax+B->rmax; q->cnt=cnt; }

/* ---- solid boxes.  A box primitive is solid: besides its surface (12 triangles, handled like any mesh) every element of the
 * other geometry is measured against the solid through one reference point -- a triangle's first vertex, a sphere's centre:
 *   d(element, solid) = dist(reference point, solid box) - radius        (0 for a point inside)
 * For a triangle outside the box the surface distance is the true one and never larger than this; for a triangle inside, the
 * surface test sees nothing and this term is 0.  For spheres it is the true signed distance by itself. */
static double point_solid_box_dist(const geom_t* X, const xf_t* Tx, const double* pw) {
  double pl[3], d[3], q[3]; v_sub(pw,Tx->t,d);
  for (int k=0;k<3;k++) pl[k]=Tx->R[k]*d[0]+Tx->R[3+k]*d[1]+Tx->R[6+k]*d[2];        /* into X's local frame */
  v_sub(pl,X->bc,d);
  double s=0;
  for (int k=0;k<3;k++) { q[k]=X->bR[k]*d[0]+X->bR[3+k]*d[1]+X->bR[6+k]*d[2];       /* onto the box axes (columns of bR) */
    double g=fabs(q[k])-X->bh[k]; if (g>0) s+=g*g; }
  return sqrt(s);
}
/* min over the elements of Y of d(element, solid X); stops early below `stop` (pass -DBL_MAX for the exact minimum) */
static double solid_min_distance(const geom_t* X, const xf_t* Tx, const geom_t* Y, const xf_t* Ty, double stop) {
  double best=DBL_MAX;
  int n=(Y->kind==G_MESH)?Y->nt:Y->np;
  for (int i=0;i<n;i++) { double pw[3], r=0;
    if (Y->kind==G_MESH) xf_apply(Ty,Y->tv+9*(size_t)i,pw); else { xf_apply(Ty,Y->pts+3*(size_t)i,pw); r=Y->rad[i]; }
    double d=point_solid_box_dist(X,Tx,pw)-r; if (d<best) { best=d; if (best<=stop) return best; } }
  return best;
}

/* a11/a13: AnyCollisionQuery::Collide / WithinDistance(tol).  Margins add to the threshold (a12). */
static int geom_pair_collide(const geom_t* A, const xf_t* Ta, const geom_t* B, const xf_t* Tb, double tol, ko_counts* cnt) {
  if (A->kind==G_EMPTY || B->kind==G_EMPTY) return 0;
  pairq_t q; pairq_init(&q,A,Ta,B,Tb,tol+A->margin+B->margin,cnt);
  if (collide_rec(&q,0,0)) return 1;
  if (A->solid && solid_min_distance(A,Ta,B,Tb,q.tol)<=q.tol) return 1;
  if (B->solid && solid_min_distance(B,Tb,A,Ta,q.tol)<=q.tol) return 1;
  return 0;
}
/* AnyCollisionQuery::Distance(0,0,bound): geometric distance minus margins; returns bound if nothing closer */
static double geom_pair_distance(const geom_t* A, const xf_t* Ta, const geom_t* B, const xf_t* Tb, double bound, ko_counts* cnt) {
  if (A->kind==G_EMPTY || B->kind==G_EMPTY) return INFINITY;
  double m=A->margin+B->margin;
  pairq_t q; pairq_init(&q,A,Ta,B,Tb,0.0,cnt);
  double best = isinf(bound)? DBL_MAX : bound+m;
  double best0=best;
  distance_rec(&q,0,0,&best);
  if (A->solid) { double d=solid_min_distance(A,Ta,B,Tb,-DBL_MAX); if (d<best) best=d; }
  if (B->solid) { double d=solid_min_distance(B,Tb,A,Ta,-DBL_MAX); if (d<best) best=d; }
  if (best>=best0) return bound;
  return best-m;
}
int ko_geom_collides(const ko_world* w, int ga, const double Ta[12], int gb, const double Tb[12]) {
  xf_t A,B; xf_from12(Ta,&A); xf_from12(Tb,&B); return geom_pair_collide(&w->geoms[ga],&A,&w->geoms[gb],&B,0.0,NULL); }
int ko_geom_within_distance(const ko_world* w, int ga, const double Ta[12], int gb, const double Tb[12], double tol) {
  xf_t A,B; xf_from12(Ta,&A); xf_from12(Tb,&B); return geom_pair_collide(&w->geoms[ga],&A,&w->geoms[gb],&B,tol,NULL); }
double ko_geom_distance(const ko_world* w, int ga, const double Ta[12], int gb, const double Tb[12], double ub) {
  xf_t A,B; xf_from12(Ta,&A); xf_from12(Tb,&B); return geom_pair_distance(&w->geoms[ga],&A,&w->geoms[gb],&B,ub,NULL); }

/* exhaustive distance over all element pairs (second, independent method for the BVH path) */
static double geom_pair_distance_brute(const geom_t* A, const xf_t* Ta, const geom_t* B, const xf_t* Tb) {
  if (A->kind==G_EMPTY || B->kind==G_EMPTY) return INFINITY;
  double best=DBL_MAX;
  int na=(A->kind==G_MESH)?A->nt:A->np, nb=(B->kind==G_MESH)?B->nt:B->np;
  for (int i=0;i<na;i++) {
    double ea[9]; double ra=0;
    if (A->kind==G_MESH) { for (int v=0;v<3;v++) xf_apply(Ta,A->tv+9*(size_t)i+3*v,ea+3*v); }
    else { xf_apply(Ta,A->pts+3*(size_t)i,ea); ra=A->rad[i]; }
    for (int j=0;j<nb;j++) {
      double eb[9]; double rb=0, d;
      if (B->kind==G_MESH) { for (int v=0;v<3;v++) xf_apply(Tb,B->tv+9*(size_t)j+3*v,eb+3*v); }
      else { xf_apply(Tb,B->pts+3*(size_t)j,eb); rb=B->rad[j]; }
      if (A->kind==G_MESH && B->kind==G_MESH) d=sqrt(tri_tri_dist2(ea,eb));
      else if (A->kind==G_MESH) d=sqrt(point_tri_dist2(eb,ea,ea+3,ea+6))-rb;
      else if (B->kind==G_MESH) d=sqrt(point_tri_dist2(ea,eb,eb+3,eb+6))-ra;
      else d=sqrt(v_dist2(ea,eb))-ra-rb;
      if (d<best) best=d;
    }
  }
  if (A->solid) { double d=solid_min_distance(A,Ta,B,Tb,-DBL_MAX); if (d<best) best=d; }
  if (B->solid) { double d=solid_min_distance(B,Tb,A,Ta,-DBL_MAX); if (d<best) best=d; }
  return best-A->margin-B->margin;
}
double ko_geom_distance_brute(const ko_world* w, int ga, const double Ta[12], int gb, const double Tb[12]) {
  xf_t A,B; xf_from12(Ta,&A); xf_from12(Tb,&B); return geom_pair_distance_brute(&w->geoms[ga],&A,&w->geoms[gb],&B); }

/* ------------------------------------------------------------------ IsFeasible */
typedef struct { const geom_t* g; xf_t T; int id; double lo[3], hi[3]; } active_t;

static int id_robot(const ko_world* w) { return w->nterr+w->nobj; }
static int id_link(const ko_world* w, int j) { return w->nterr+w->nobj+1+j; }
static inline int mask_en(const ko_world* w, int a, int b) { return w->mask[(size_t)a*w->nids+b]; }

/* GetGeometries (PlannerSettings.cpp:214-239) for the robot and for {terrains, rigid objects} */
static int gather_links(const ko_world* w, const xf_t* T, active_t* out) {
  int n=0;
  for (int j=0;j<w->L;j++) { int g=w->linkgeom[j]; if (geom_empty(w,g)) continue;
    out[n].g=&w->geoms[g]; out[n].T=T[j]; out[n].id=id_link(w,j); n++; }
  return n;
}
static int gather_env(const ko_world* w, active_t* out) {
  int n=0; xf_t I; xf_identity(&I);
  for (int i=0;i<w->nterr;i++) { int g=w->terrains[i]; if (geom_empty(w,g)) continue; out[n].g=&w->geoms[g]; out[n].T=I; out[n].id=i; n++; }
  for (int i=0;i<w->nobj;i++) { int g=w->objects[i]; if (geom_empty(w,g)) continue; out[n].g=&w->geoms[g]; out[n].T=w->objT[i]; out[n].id=w->nterr+i; n++; }
  return n;
}
static void compute_bbs(active_t* a, int n, double tol) {
  for (int i=0;i<n;i++) { box_world_aabb(a[i].g->lo,a[i].g->hi,&a[i].T,a[i].g->margin,a[i].lo,a[i].hi);
    for (int k=0;k<3;k++) { a[i].lo[k]-=0.5*tol; a[i].hi[k]+=0.5*tol; } }
}
/* env check, PlannerSettings.cpp:269-331 (two group-AABB quick rejects, then pair loop) */
static int check_env(const ko_world* w, active_t* s1, int n1, active_t* s2, int n2, double tol, int32_t* pair, ko_counts* cnt) {
  compute_bbs(s1,n1,tol); compute_bbs(s2,n2,tol);
  double lo[3]={DBL_MAX,DBL_MAX,DBL_MAX}, hi[3]={-DBL_MAX,-DBL_MAX,-DBL_MAX};
  for (int i=0;i<n1;i++) for (int k=0;k<3;k++) { if (s1[i].lo[k]<lo[k]) lo[k]=s1[i].lo[k]; if (s1[i].hi[k]>hi[k]) hi[k]=s1[i].hi[k]; }
  for (int i=0;i<n2;i++) { if (cnt) cnt->n_box++; if (!aabb_overlap(s2[i].lo,s2[i].hi,lo,hi)) { s2[i]=s2[n2-1]; n2--; i--; } }
  for (int k=0;k<3;k++) { lo[k]=DBL_MAX; hi[k]=-DBL_MAX; }
  for (int i=0;i<n2;i++) for (int k=0;k<3;k++) { if (s2[i].lo[k]<lo[k]) lo[k]=s2[i].lo[k]; if (s2[i].hi[k]>hi[k]) hi[k]=s2[i].hi[k]; }
  for (int i=0;i<n1;i++) { if (cnt) cnt->n_box++; if (!aabb_overlap(s1[i].lo,s1[i].hi,lo,hi)) { s1[i]=s1[n1-1]; n1--; i--; } }
  for (int i=0;i<n1;i++) for (int j=0;j<n2;j++) {
    if (mask_en(w,s1[i].id,s2[j].id) || mask_en(w,s2[j].id,s1[i].id)) {
      if (cnt) cnt->n_box++;
      if (aabb_overlap(s1[i].lo,s1[i].hi,s2[j].lo,s2[j].hi))
        if (geom_pair_collide(s1[i].g,&s1[i].T,s2[j].g,&s2[j].T,tol,cnt)) { if (pair) { pair[0]=s1[i].id; pair[1]=s2[j].id; } return 1; }
    } }
  return 0;
}
/* self check, PlannerSettings.cpp:241-267 (second mask term indexes the diagonal: always false for links) */
static int check_self(const ko_world* w, active_t* s, int n, double tol, int32_t* pair, ko_counts* cnt) {
  compute_bbs(s,n,tol);
  for (int i=0;i<n;i++) for (int j=i+1;j<n;j++) {
    if (mask_en(w,s[i].id,s[j].id) || mask_en(w,s[i].id,s[i].id)) {
      if (cnt) cnt->n_box++;
      if (aabb_overlap(s[i].lo,s[i].hi,s[j].lo,s[j].hi))
        if (geom_pair_collide(s[i].g,&s[i].T,s[j].g,&s[j].T,tol,cnt)) { if (pair) { pair[0]=s[i].id; pair[1]=s[j].id; } return 1; }
    } }
  return 0;
}
static int check_collision_free(const ko_world* w, const double* q, int32_t* pair, ko_counts* cnt) {
  int L=w->L;
  int big = (L+w->nterr+w->nobj) > 2048;
  size_t b1=sizeof(xf_t)*L, b2=sizeof(active_t)*(L+1), b3=sizeof(active_t)*(w->nterr+w->nobj+1);
  xf_t* T=(xf_t*)(big?malloc(b1):alloca(b1)); fk_links(w,q,T);
  active_t* s1=(active_t*)(big?malloc(b2):alloca(b2));
  active_t* s2=(active_t*)(big?malloc(b3):alloca(b3));
  int n1=gather_links(w,T,s1), n2=gather_env(w,s2), hit;
  hit=check_env(w,s1,n1,s2,n2,0.0,pair,cnt);
  if (!hit) { n1=gather_links(w,T,s1); hit=check_self(w,s1,n1,0.0,pair,cnt); }
  if (big) { free(T); free(s1); free(s2); }
  return !hit;
}
int ko_feasible(const ko_world* w, const double* q, int32_t* pair, ko_counts* cnt) {
  if (pair) { pair[0]=-1; pair[1]=-1; }
  if (!ko_check_joint_limits(w,q)) return 0;
  return check_collision_free(w,q,pair,cnt);
}
/* all element pairs of all enabled geometry pairs; no BVH, no AABB reject */
int ko_feasible_brute(const ko_world* w, const double* q) {
  if (!ko_check_joint_limits(w,q)) return 0;
  int L=w->L; xf_t* T=(xf_t*)malloc(sizeof(xf_t)*L); fk_links(w,q,T);
  active_t* s1=(active_t*)malloc(sizeof(active_t)*(L+1));
  active_t* s2=(active_t*)malloc(sizeof(active_t)*(w->nterr+w->nobj+1));
  int n1=gather_links(w,T,s1), n2=gather_env(w,s2), hit=0;
  for (int i=0;i<n1&&!hit;i++) for (int j=0;j<n2&&!hit;j++)
    if (mask_en(w,s1[i].id,s2[j].id)||mask_en(w,s2[j].id,s1[i].id))
      hit = geom_pair_distance_brute(s1[i].g,&s1[i].T,s2[j].g,&s2[j].T) <= 0.0;
  for (int i=0;i<n1&&!hit;i++) for (int j=i+1;j<n1&&!hit;j++)
    if (mask_en(w,s1[i].id,s1[j].id))
      hit = geom_pair_distance_brute(s1[i].g,&s1[i].T,s1[j].g,&s1[j].T) <= 0.0;
  free(T); free(s1); free(s2);
  return !hit;
}
/* ------------------------------------------------------------------ contact depth (test support for the 1e-6 m band)
 * "Boolean results bit-exact outside a 1e-6 m margin band" needs a two-sided definition of the band.  A configuration the oracle
 * calls FREE is inside the band when its clearance (ko_distance) is <= 1e-6.  A configuration the oracle calls COLLIDING is inside
 * the band when its deepest contact is <= 1e-6: the largest, over all enabled geometry pairs and all element pairs within the
 * pair's threshold, of (threshold - signed distance), where the signed distance of two intersecting triangles is minus their
 * penetration depth -- the smallest translation that separates them = the smallest overlap of their projections on the 11
 * candidate axes (two face normals, nine edge cross products: the face normals of the Minkowski difference of two triangles). */
static double tri_tri_depth(const double* A, const double* B) {
  double ax[11][3]; int na=0; double e[6][3];
  for (int i=0;i<3;i++) { v_sub(A+3*((i+1)%3),A+3*i,e[i]); v_sub(B+3*((i+1)%3),B+3*i,e[3+i]); }
  v_cross(e[0],e[1],ax[na++]); v_cross(e[3],e[4],ax[na++]);
  for (int i=0;i<3;i++) for (int j=0;j<3;j++) v_cross(e[i],e[3+j],ax[na++]);
  double best=DBL_MAX;
  for (int k=0;k<na;k++) { double n2=v_dot(ax[k],ax[k]); if (!(n2>1e-300)) continue; double inv=1.0/sqrt(n2);
    double amin=DBL_MAX,amax=-DBL_MAX,bmin=DBL_MAX,bmax=-DBL_MAX;
    for (int v=0;v<3;v++) { double pa=v_dot(A+3*v,ax[k])*inv, pb=v_dot(B+3*v,ax[k])*inv;
      amin=fmin(amin,pa); amax=fmax(amax,pa); bmin=fmin(bmin,pb); bmax=fmax(bmax,pb); }
    double o=fmin(amax-bmin,bmax-amin); if (o<0) o=0; if (o<best) best=o; }
  return best==DBL_MAX?0.0:best;
}
static void depth_leaf(pairq_t* q, const node_t* a, const node_t* b, double* worst) {
  const geom_t *A=q->A, *B=q->B; double d;
  if (A->kind==G_MESH && B->kind==G_MESH) {
    double ta[9], tb[9];
    for (int v=0;v<3;v++) { xf_apply(&q->Ta,A->tv+9*(size_t)a->first+3*v,ta+3*v); xf_apply(&q->Tb,B->tv+9*(size_t)b->first+3*v,tb+3*v); }
    if (ko_tri_tri_intersect(ta,tb)) d=-tri_tri_depth(ta,tb); else d=sqrt(tri_tri_dist2(ta,tb));
    if (d<=q->tol && q->tol-d>*worst) *worst=q->tol-d;
    return;
  }
  if (A->kind==G_MESH || B->kind==G_MESH) {
    const geom_t* M = A->kind==G_MESH?A:B; const geom_t* C = A->kind==G_MESH?B:A;
    const node_t* mn = A->kind==G_MESH?a:b; const node_t* cn = A->kind==G_MESH?b:a;
    const xf_t* Tm = A->kind==G_MESH?&q->Ta:&q->Tb; const xf_t* Tc = A->kind==G_MESH?&q->Tb:&q->Ta;
    double t[9]; for (int v=0;v<3;v++) xf_apply(Tm,M->tv+9*(size_t)mn->first+3*v,t+3*v);
    for (int i=cn->first;i<cn->first+cn->count;i++) { double p[3]; xf_apply(Tc,C->pts+3*(size_t)i,p);
      d=sqrt(point_tri_dist2(p,t,t+3,t+6))-C->rad[i]; if (d<=q->tol && q->tol-d>*worst) *worst=q->tol-d; }
    return;
  }
  for (int i=a->first;i<a->first+a->count;i++) { double p[3]; xf_apply(&q->Ta,A->pts+3*(size_t)i,p);
    for (int j=b->first;j<b->first+b->count;j++) { double s[3]; xf_apply(&q->Tb,B->pts+3*(size_t)j,s);
      d=sqrt(v_dist2(p,s))-A->rad[i]-B->rad[j]; if (d<=q->tol && q->tol-d>*worst) *worst=q->tol-d; } }
}
static void depth_rec(pairq_t* q, int ia, int ib, double* worst) {
  const node_t* a=&q->A->nodes[ia]; const node_t* b=&q->B->nodes[ib];
  if (!obb_overlap(a,b,&q->Tab,q->tol)) return;
  int la=a->left<0, lb=b->left<0;
  if (la && lb) { depth_leaf(q,a,b,worst); return; }
  if (lb || (!la && node_size2(a)>=node_size2(b))) { depth_rec(q,a->left,ib,worst); depth_rec(q,node_right(q->A,ia),ib,worst); }
  else { depth_rec(q,ia,b->left,worst); depth_rec(q,ia,node_right(q->B,ib),worst); }
}
static void geom_pair_depth(const geom_t* A, const xf_t* Ta, const geom_t* B, const xf_t* Tb, double* worst) {
  if (A->kind==G_EMPTY || B->kind==G_EMPTY) return;
  pairq_t q; pairq_init(&q,A,Ta,B,Tb,A->margin+B->margin,NULL);
  depth_rec(&q,0,0,worst);
  /* solids: an element reference point inside the box has no measurable depth here -> reported as 1 m (never inside the band) */
  for (int side=0;side<2;side++) { const geom_t* X=side?B:A; const geom_t* Y=side?A:B; const xf_t* Tx=side?Tb:Ta; const xf_t* Ty=side?Ta:Tb;
    if (!X->solid) continue;
    int n=(Y->kind==G_MESH)?Y->nt:Y->np;
    for (int i=0;i<n;i++) { double pw[3], r=0;
      if (Y->kind==G_MESH) xf_apply(Ty,Y->tv+9*(size_t)i,pw); else { xf_apply(Ty,Y->pts+3*(size_t)i,pw); r=Y->rad[i]; }
      double d0=point_solid_box_dist(X,Tx,pw), d=d0-r;
      if (d<=q.tol) { double dep = d0==0.0 ? 1.0 : q.tol-d; if (dep>*worst) *worst=dep; } } }
}
/* deepest contact of configuration q (metres); -1 when nothing is in contact; joint limits are not looked at */
double ko_penetration(const ko_world* w, const double* q, int include_self) {
  int L=w->L; xf_t* T=(xf_t*)malloc(sizeof(xf_t)*L); fk_links(w,q,T);
  active_t* s1=(active_t*)malloc(sizeof(active_t)*(L+1));
  active_t* s2=(active_t*)malloc(sizeof(active_t)*(w->nterr+w->nobj+1));
  int n1=gather_links(w,T,s1), n2=gather_env(w,s2);
  double worst=-1.0;
  for (int i=0;i<n1;i++) for (int j=0;j<n2;j++)
    if (mask_en(w,s1[i].id,s2[j].id)||mask_en(w,s2[j].id,s1[i].id)) geom_pair_depth(s1[i].g,&s1[i].T,s2[j].g,&s2[j].T,&worst);
  if (include_self) for (int i=0;i<n1;i++) for (int j=i+1;j<n1;j++)
    if (mask_en(w,s1[i].id,s1[j].id)) geom_pair_depth(s1[i].g,&s1[i].T,s1[j].g,&s1[j].T,&worst);
  free(T); free(s1); free(s2);
  return worst;
}
double ko_tri_tri_depth(const double a[9], const double b[9]) { return tri_tri_depth(a,b); }
/* the same for one geometry pair at explicit transforms with threshold margins + tol */
double ko_geom_penetration(const ko_world* w, int ga, const double Ta[12], int gb, const double Tb[12], double tol) {
  xf_t A,B; xf_from12(Ta,&A); xf_from12(Tb,&B);
  geom_t a=w->geoms[ga], b=w->geoms[gb]; a.margin+=tol;      /* shallow copies: the threshold is margin_A + margin_B + tol */
  double worst=-1.0; geom_pair_depth(&a,&A,&b,&B,&worst); return worst; }

/* ------------------------------------------------------------------ ray casting
 * WorldModel::RayCast / RayCastIgnore (reference Cpp/Modeling/World.cpp:465-588): the closest hit over the robot's links at q, the
 * rigid objects and the terrains, visited in that order with a strict '<' (a tie keeps the earlier body); each body answers through
 * AnyCollisionGeometry3D::RayCast(ray, &dist) and the world point is source + dist * direction.  Geometry3D::rayCast / rayCast_ext
 * (Python/klampt/src/geometry.cpp:1821-1852) is the one-geometry form.  The per-geometry arithmetic lives in KrisLibrary (absent:
 * parity unpinned); restated here from its published behaviour: a triangle mesh reports the nearest two-sided ray / triangle
 * intersection (watertight form, see ray_tri) and takes its collision margin off the distance; a point cloud is the union of spheres of radius (point radius +
 * margin) and reports where the ray enters the first one (nothing when that radius is 0).  The direction is normalised on entry, so
 * dist is a length. */
static __thread ko_counts* ray_cnt = NULL;      /* set by ko_raycast_counts: box / triangle / sphere tests of the calling thread */
/* Ray vs triangle, two-sided, watertight: the triangle is sheared into the ray's frame (largest direction component along z) and
 * tested with three 2-D edge functions whose sign is exact -- each a difference of two products evaluated with its rounding error
 * (fma) -- and antisymmetric in the edge's two vertices, so a ray through an edge or a vertex shared by two triangles always hits
 * at least one of them and never slips between (Woop, Benthin, Wald 2013).  KrisLibrary's own test (absent) is the plain plane /
 * barycentric form; results differ only for rays that pass exactly through an edge.  t = distance along d. */
typedef struct { int kx, ky, kz; double Sx, Sy, Sz; } rayshear_t;
static void ray_shear(const double* d, rayshear_t* r) {
  int kz=0; if (fabs(d[1])>fabs(d[kz])) kz=1; if (fabs(d[2])>fabs(d[kz])) kz=2;
  int kx=(kz+1)%3, ky=(kx+1)%3;
  if (d[kz]<0) { int t=kx; kx=ky; ky=t; }
  r->kx=kx; r->ky=ky; r->kz=kz; r->Sx=d[kx]/d[kz]; r->Sy=d[ky]/d[kz]; r->Sz=1.0/d[kz];
}
static inline double diff_of_products(double a, double b, double c, double d) {   /* a*b - c*d with an exact sign */
  double p1=a*b, e1=fma(a,b,-p1), p2=c*d, e2=fma(c,d,-p2);
  return (p1-p2)+(e1-e2);
}
static int ray_tri(const double* s, const rayshear_t* r, const double* a, const double* b, const double* c, double* t) {
  if (ray_cnt) ray_cnt->n_tri++;
  const int kx=r->kx, ky=r->ky, kz=r->kz;
  const double Az=a[kz]-s[kz], Bz=b[kz]-s[kz], Cz=c[kz]-s[kz];
  const double Ax=(a[kx]-s[kx])-r->Sx*Az, Ay=(a[ky]-s[ky])-r->Sy*Az;
  const double Bx=(b[kx]-s[kx])-r->Sx*Bz, By=(b[ky]-s[ky])-r->Sy*Bz;
  const double Cx=(c[kx]-s[kx])-r->Sx*Cz, Cy=(c[ky]-s[ky])-r->Sy*Cz;
  const double U=diff_of_products(Cx,By,Cy,Bx), V=diff_of_products(Ax,Cy,Ay,Cx), W=diff_of_products(Bx,Ay,By,Ax);
  if ((U<0||V<0||W<0) && (U>0||V>0||W>0)) return 0;
  const double det=U+V+W;
  if (!(det!=0)) return 0;                                   /* edge-on, or no area (segment triangles) */
  const double T=(U*(r->Sz*Az)+V*(r->Sz*Bz))+W*(r->Sz*Cz);
  if ((det<0 && T>0) || (det>0 && T<0)) return 0;            /* behind the source */
  const double tt=T/det;
  if (!(tt>=0)) return 0;
  *t=tt; return 1;
}
static int ray_sphere(const double* s, const double* d, const double* c, double r, double* t) {   /* |d| = 1 */
  if (ray_cnt) ray_cnt->n_pt++;
  if (!(r>0)) return 0;
  double m[3]; v_sub(s,c,m); double b=v_dot(m,d), cc=v_dot(m,m)-r*r;
  if (cc<=0) { *t=0; return 1; }                              /* the source is inside */
  if (b>0) return 0;
  double disc=b*b-cc; if (disc<0) return 0;
  *t=-b-sqrt(disc); if (*t<0) *t=0; return 1;
}
static int ray_box(const double* s, const double* d, const double* lo, const double* hi, double pad, double tmax, double* tnear) {
  double t0=0, t1=tmax;
  if (ray_cnt) ray_cnt->n_node++;
  for (int k=0;k<3;k++) {
    if (d[k]==0) { if (s[k]<lo[k]-pad || s[k]>hi[k]+pad) return 0; continue; }
    double a=(lo[k]-pad-s[k])/d[k], b=(hi[k]+pad-s[k])/d[k];
    double mn=a<b?a:b, mx=a<b?b:a;
    if (mn>t0) t0=mn;
    if (mx<t1) t1=mx;
  }
  *tnear=t0; return t0<=t1;
}
/* nearest hit of the local-frame ray with geometry g: t (before the margin is taken off) and the element in the caller's order */
static int geom_ray_local(const geom_t* g, const double* s, const double* d, double tmax, double* tbest, int* elem, int brute) {
  int hit=0; *tbest=tmax;
  rayshear_t sh; ray_shear(d,&sh);
  if (g->kind==G_EMPTY || g->nnodes<=0) return 0;
  if (brute) {
    int n=(g->kind==G_MESH)?g->nt:g->np;
    for (int i=0;i<n;i++) { double t; int h=(g->kind==G_MESH) ? ray_tri(s,&sh,g->tv+9*(size_t)i,g->tv+9*(size_t)i+3,g->tv+9*(size_t)i+6,&t)
                                                              : ray_sphere(s,d,g->pts+3*(size_t)i,g->rad[i]+g->margin,&t);
      if (h && (!hit || t<*tbest || (t==*tbest && g->perm[i]<*elem))) { *tbest=t; *elem=g->perm[i]; hit=1; } }
    return hit;
  }
  double pad=(g->kind==G_MESH)?0.0:g->margin;                 /* node boxes hold the point radii, not the margin */
  int stack[128], sp=0; stack[sp++]=0;
  while (sp>0) {
    int i=stack[--sp]; const node_t* n=&g->nodes[i]; double tn;
    if (!ray_box(s,d,n->lo,n->hi,pad+1e-12,*tbest,&tn)) continue;
    if (n->left<0) {
      for (int e=n->first;e<n->first+n->count;e++) { double t; int h=(g->kind==G_MESH) ? ray_tri(s,&sh,g->tv+9*(size_t)e,g->tv+9*(size_t)e+3,g->tv+9*(size_t)e+6,&t)
                                                                                    : ray_sphere(s,d,g->pts+3*(size_t)e,g->rad[e]+g->margin,&t);
        if (h && (!hit || t<*tbest || (t==*tbest && g->perm[e]<*elem))) { *tbest=t; *elem=g->perm[e]; hit=1; } }
    } else if (sp+2<=128) { stack[sp++]=n->left; stack[sp++]=n->right; }
  }
  return hit;
}
static int geom_raycast(const geom_t* g, const xf_t* T, const double* s, const double* d, double* dist, int* elem, int brute) {
  double sl[3],dl[3],m[3]; v_sub(s,T->t,m);
  for (int k=0;k<3;k++) { sl[k]=T->R[k]*m[0]+T->R[3+k]*m[1]+T->R[6+k]*m[2]; dl[k]=T->R[k]*d[0]+T->R[3+k]*d[1]+T->R[6+k]*d[2]; }   /* R^T */
  double t; if (!geom_ray_local(g,sl,dl,DBL_MAX,&t,elem,brute)) return 0;
  *dist = (g->kind==G_MESH) ? t-g->margin : t;
  return 1;
}
static int ray_unit(const double* d, double* u) { double n=sqrt(v_dot(d,d)); if (!(n>0) || !isfinite(n)) return 0; u[0]=d[0]/n; u[1]=d[1]/n; u[2]=d[2]/n; return 1; }
int ko_geom_raycast(const ko_world* w, int g, const double T12[12], const double s[3], const double d[3], double* dist, int32_t* elem, int brute) {
  xf_t T; xf_from12(T12,&T); double u[3]; int el=-1; *dist=INFINITY; if (elem) *elem=-1;
  if (g<0||g>=w->ngeoms||!ray_unit(d,u)) return 0;
  if (!geom_raycast(&w->geoms[g],&T,s,u,dist,&el,brute)) { *dist=INFINITY; return 0; }
  if (elem) *elem=el;
  return 1;
}
int ko_raycast(const ko_world* w, const double* q, const double s[3], const double d[3], const uint8_t* ignore_ids, double* dist, int32_t* elem) {
  double u[3]; int best=-1, bel=-1; double bd=INFINITY;
  if (ray_unit(d,u)) {
    if (q && w->L) {
      xf_t* T=(xf_t*)malloc(sizeof(xf_t)*w->L); fk_links(w,q,T);
      for (int j=0;j<w->L;j++) { int g=w->linkgeom[j]; if (geom_empty(w,g) || (ignore_ids && ignore_ids[id_link(w,j)])) continue;
        double dd; int el; if (geom_raycast(&w->geoms[g],&T[j],s,u,&dd,&el,0) && dd<bd) { bd=dd; best=id_link(w,j); bel=el; } }
      free(T);
    }
    for (int i=0;i<w->nobj;i++) { int g=w->objects[i]; if (geom_empty(w,g) || (ignore_ids && ignore_ids[w->nterr+i])) continue;
      double dd; int el; if (geom_raycast(&w->geoms[g],&w->objT[i],s,u,&dd,&el,0) && dd<bd) { bd=dd; best=w->nterr+i; bel=el; } }
    xf_t I; xf_identity(&I);
    for (int i=0;i<w->nterr;i++) { int g=w->terrains[i]; if (geom_empty(w,g) || (ignore_ids && ignore_ids[i])) continue;
      double dd; int el; if (geom_raycast(&w->geoms[g],&I,s,u,&dd,&el,0) && dd<bd) { bd=dd; best=i; bel=el; } }
  }
  *dist=bd; if (elem) *elem=bel; return best;
}
/* traversal counts of the per-body loop above, summed over N rays (single thread): n_node = box tests, n_tri / n_pt = element tests */
void ko_raycast_counts(const ko_world* w, const double* q, const double* rays, int64_t N, ko_counts* total) {
  memset(total,0,sizeof(*total)); ray_cnt=total;
  for (int64_t i=0;i<N;i++) { double dd; int32_t el; ko_raycast(w,q,rays+6*i,rays+6*i+3,NULL,&dd,&el); }
  ray_cnt=NULL;
}
void ko_raycast_batch(const ko_world* w, const double* q, const double* rays, int64_t N, const uint8_t* ignore_ids, int32_t* ids, double* dist, int32_t* elem, int nthreads) {
  if (nthreads<=0) nthreads=ko_max_threads();
  #pragma omp parallel for schedule(dynamic,256) num_threads(nthreads)
  for (int64_t i=0;i<N;i++) { int32_t el; ids[i]=ko_raycast(w,q,rays+6*i,rays+6*i+3,ignore_ids,dist+i,&el); if (elem) elem[i]=el; }
}

int ko_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void ko_feasible_batch(const ko_world* w, const double* Q, int64_t N, uint8_t* out, int32_t* first_pair, ko_counts* cpc, int nthreads) {
  int L=w->L;
  if (nthreads<=0) nthreads=ko_max_threads();
  #pragma omp parallel for schedule(dynamic,64) num_threads(nthreads)
  for (int64_t c=0;c<N;c++) {
    int32_t pr[2]; ko_counts cnt={0,0,0,0};
    out[c]=(uint8_t)ko_feasible(w,Q+c*L,pr,cpc?&cnt:NULL);
    if (first_pair) { first_pair[2*c]=pr[0]; first_pair[2*c+1]=pr[1]; }
    if (cpc) cpc[c]=cnt;
  }
}

/* ------------------------------------------------------------------ edges */
static double angle_normalize(double a) { a=fmod(a,2*M_PI); if (a<0) a+=2*M_PI; return a; }
static double angle_diff(double a, double b) { /* signed CCW difference a-b in (-pi,pi] */
  double d=a-b; if (d>M_PI) return d-2*M_PI; if (d<-M_PI) return d+2*M_PI; return d; }
/* ---- SO(3) helpers for Floating / BallAndSocket joints.  KrisLibrary's EulerAngleRotation / interpolateRotation /
 * AngleAxisRotation are absent; these restate their documented meaning: getMatrixZYX(a,b,c) = Rz(a) Ry(b) Rx(c) (the
 * reference asserts the three links turn about z, y, x in that order, Interpolate.cpp:24-26), interpolateRotation = the
 * SO(3) geodesic Rx exp(u log(Rx^T Ry)), AngleAxisRotation::angle = acos((tr - 1) / 2).  [unverified here: branch
 * thresholds of the library's own log / Euler extraction near angle pi and gimbal lock] */
static void euler_zyx_to_matrix(double a, double b, double c, double R[9]) {
  double ca=cos(a), sa=sin(a), cb=cos(b), sb=sin(b), cc=cos(c), sc=sin(c);
  R[0]=ca*cb; R[1]=ca*sb*sc-sa*cc; R[2]=ca*sb*cc+sa*sc;
  R[3]=sa*cb; R[4]=sa*sb*sc+ca*cc; R[5]=sa*sb*cc-ca*sc;
  R[6]=-sb;   R[7]=cb*sc;          R[8]=cb*cc;
}
static void matrix_to_euler_zyx(const double R[9], double* a, double* b, double* c) {
  double sb=-R[6]; if (sb>1) sb=1; if (sb<-1) sb=-1;
  *b=asin(sb);
  if (fabs(R[6])<1.0-1e-12) { *a=atan2(R[3],R[0]); *c=atan2(R[7],R[8]); }
  else { *c=0; *a=atan2(-R[1],R[4]); }          /* gimbal lock: put the whole turn into the z angle */
}
static void mat3_mul_tA(const double A[9], const double B[9], double C[9]) {   /* C = A^T B */
  for (int i=0;i<3;i++) for (int j=0;j<3;j++) C[3*i+j]=A[i]*B[j]+A[3+i]*B[3+j]+A[6+i]*B[6+j]; }
static void mat3_mul(const double A[9], const double B[9], double C[9]) {
  for (int i=0;i<3;i++) for (int j=0;j<3;j++) C[3*i+j]=A[3*i]*B[j]+A[3*i+1]*B[3+j]+A[3*i+2]*B[6+j]; }
static double so3_angle(const double R[9]) { double c=0.5*(R[0]+R[4]+R[8]-1.0); if (c>1) c=1; if (c<-1) c=-1; return acos(c); }
static void so3_log(const double R[9], double w[3]) {
  double th=so3_angle(R);
  double v[3]={R[7]-R[5], R[2]-R[6], R[3]-R[1]};
  if (th<1e-9) { for (int k=0;k<3;k++) w[k]=0.5*v[k]; return; }
  if (M_PI-th<1e-6) {   /* near a half turn: axis from the diagonal of (R + I) / 2, sign from the skew part */
    double ax[3]; for (int k=0;k<3;k++) { double d=0.5*(R[4*k]+1.0); ax[k]=d>0?sqrt(d):0; }
    int m=0; for (int k=1;k<3;k++) if (ax[k]>ax[m]) m=k;
    for (int k=0;k<3;k++) if (k!=m) { double s=R[3*m+k]+R[3*k+m]; if (s<0) ax[k]=-ax[k]; }
    double sg=ax[0]*v[0]+ax[1]*v[1]+ax[2]*v[2]; if (sg<0) for (int k=0;k<3;k++) ax[k]=-ax[k];
    double n=sqrt(ax[0]*ax[0]+ax[1]*ax[1]+ax[2]*ax[2]); for (int k=0;k<3;k++) w[k]=th*ax[k]/n; return; }
  double f=th/(2.0*sin(th)); for (int k=0;k<3;k++) w[k]=f*v[k];
}
static void so3_exp(const double w[3], double R[9]) {   /* Rodrigues */
  double th=sqrt(w[0]*w[0]+w[1]*w[1]+w[2]*w[2]);
  if (th<1e-12) { R[0]=1; R[1]=-w[2]; R[2]=w[1]; R[3]=w[2]; R[4]=1; R[5]=-w[0]; R[6]=-w[1]; R[7]=w[0]; R[8]=1; return; }
  double x=w[0]/th, y=w[1]/th, z=w[2]/th, c=cos(th), s=sin(th), v=1.0-c;
  R[0]=c+v*x*x;   R[1]=v*x*y-s*z; R[2]=v*x*z+s*y;
  R[3]=v*y*x+s*z; R[4]=c+v*y*y;   R[5]=v*y*z-s*x;
  R[6]=v*z*x-s*y; R[7]=v*z*y+s*x; R[8]=c+v*z*z;
}
static double euler_zyx_angle_between(const double* ea, const double* eb) {   /* angle of Ra Rb^T */
  double Ra[9], Rb[9], D[9]; euler_zyx_to_matrix(ea[0],ea[1],ea[2],Ra); euler_zyx_to_matrix(eb[0],eb[1],eb[2],Rb);
  for (int i=0;i<3;i++) for (int j=0;j<3;j++) D[3*i+j]=Ra[3*i]*Rb[3*j]+Ra[3*i+1]*Rb[3*j+1]+Ra[3*i+2]*Rb[3*j+2];
  return so3_angle(D); }
static void euler_zyx_interp(const double* ea, const double* eb, double u, double* out) {
  double Ra[9], Rb[9], D[9], w[3], E[9], Ru[9];
  euler_zyx_to_matrix(ea[0],ea[1],ea[2],Ra); euler_zyx_to_matrix(eb[0],eb[1],eb[2],Rb);
  mat3_mul_tA(Ra,Rb,D); so3_log(D,w); for (int k=0;k<3;k++) w[k]*=u; so3_exp(w,E); mat3_mul(Ra,E,Ru);
  matrix_to_euler_zyx(Ru,&out[0],&out[1],&out[2]); }

/* RobotCSpace::Distance -> Klampt::Distance, Interpolate.cpp:208-343, norm=2 (RobotCSpace.cpp:48), floatingRotationWeight=1
 * (RobotCSpace.cpp:50).  Weld joints are skipped; Normal joints collect (a-b) (times weight); Spin uses AngleDiff; Floating
 * collects the three translations and the geodesic angle; BallAndSocket the geodesic angle; FloatingPlanar contributes
 * nothing (the weighted overload's default branch, :338-340; the unweighted one aborts on it, :270). */
double ko_cspace_distance(const ko_world* w, const double* a, const double* b, const double* weights) {
  double s=0;
  for (int i=0;i<w->nj;i++) { int k=w->jlink[i]; double wt=weights?weights[i]:1.0, d; int idx[6];
    switch (w->jtype[i]) {
      case KO_JOINT_WELD: continue;
      case KO_JOINT_NORMAL: d=a[k]-b[k]; break;
      case KO_JOINT_SPIN: d=angle_diff(angle_normalize(a[k]),angle_normalize(b[k])); break;
      case KO_JOINT_FLOATING: { joint_indices(w,i,idx);
        for (int t=0;t<3;t++) { double dt=a[idx[t]]-b[idx[t]]; s+=wt*dt*dt; }
        double ea[3]={a[idx[3]],a[idx[4]],a[idx[5]]}, eb[3]={b[idx[3]],b[idx[4]],b[idx[5]]};
        d=euler_zyx_angle_between(ea,eb); break; }
      case KO_JOINT_BALLANDSOCKET: { joint_indices(w,i,idx);
        double ea[3]={a[idx[0]],a[idx[1]],a[idx[2]]}, eb[3]={b[idx[0]],b[idx[1]],b[idx[2]]};
        d=euler_zyx_angle_between(ea,eb); break; }
      default: continue; }
    s+=wt*d*d; }   /* NormAccumulator<Real>(2).collect(x,w): sum w*x^2, then sqrt */
  return sqrt(s);
}
/* Klampt::Interpolate, Interpolate.cpp:10-71: out = x*(1-u); out += y*u; Spin and FloatingPlanar angles use the shortest arc;
 * the Euler-ZYX triplet of Floating / BallAndSocket joints follows the SO(3) geodesic.  (The reference's BallAndSocket branch
 * writes the result through indices[0], [2], [3] of a 3-element vector, :50 -- out of range; the intended [0], [1], [2] is used.) */
void ko_interpolate(const ko_world* w, const double* a, const double* b, double u, double* out) {
  for (int k=0;k<w->L;k++) { out[k]=a[k]*(1.0-u); out[k]+=b[k]*u; }
  for (int i=0;i<w->nj;i++) { int t=w->jtype[i], idx[6];
    if (t==KO_JOINT_SPIN || t==KO_JOINT_FLOATINGPLANAR) {
      int k=w->jlink[i]; if (t==KO_JOINT_FLOATINGPLANAR) { joint_indices(w,i,idx); k=idx[2]; }
      double x=angle_normalize(a[k]), y=angle_normalize(b[k]); double d=angle_diff(y,x);
      out[k]=angle_normalize(x+u*d); }
    else if (t==KO_JOINT_FLOATING || t==KO_JOINT_BALLANDSOCKET) {
      joint_indices(w,i,idx); int o=t==KO_JOINT_FLOATING?3:0;
      double ea[3]={a[idx[o]],a[idx[o+1]],a[idx[o+2]]}, eb[3]={b[idx[o]],b[idx[o+1]],b[idx[o+2]]}, eu[3];
      euler_zyx_interp(ea,eb,u,eu);
      out[idx[o]]=eu[0]; out[idx[o+1]]=eu[1]; out[idx[o+2]]=eu[2]; }
  }
}
/* EpsilonEdgeChecker::IsVisible: bisect until segment length <= eps, midpoints in
 * coarse-to-fine order, endpoints not re-checked, false at the first infeasible midpoint. */
int ko_edge_visible(const ko_world* w, const double* a, const double* b, double eps, const double* weights, int32_t* nchecks, ko_counts* cnt) {
  double len=ko_cspace_distance(w,a,b,weights);
  double* m=(double*)malloc(sizeof(double)*w->L);
  int32_t n=0; int vis=1; long segs=1;
  while (len>eps && vis) {
    segs*=2; len*=0.5;
    for (long k=1;k<segs;k+=2) {
      ko_interpolate(w,a,b,(double)k/(double)segs,m);
      n++;
      if (!ko_feasible(w,m,NULL,cnt)) { vis=0; break; }
    }
  }
  free(m); if (nchecks) *nchecks=n; return vis;
}
void ko_edges_visible_batch(const ko_world* w, const double* A, const double* B, int64_t N, double eps,
                            const double* weights, uint8_t* out, int32_t* nchecks, int nthreads) {
  int L=w->L; if (nthreads<=0) nthreads=ko_max_threads();
  #pragma omp parallel for schedule(dynamic,8) num_threads(nthreads)
  for (int64_t e=0;e<N;e++) { int32_t n; out[e]=(uint8_t)ko_edge_visible(w,A+e*L,B+e*L,eps,weights,&n,NULL); if (nchecks) nchecks[e]=n; }
}

/* ------------------------------------------------------------------ distance */
static double aabb_dist(const double* alo, const double* ahi, const double* blo, const double* bhi) {
  double s=0; for (int k=0;k<3;k++) { double g=fmax(alo[k]-bhi[k],blo[k]-ahi[k]); if (g>0) s+=g*g; } return sqrt(s); }
/* WorldPlannerSettings::DistanceLowerBound with eps=0 (PlannerSettings.cpp:570-620): min over enabled pairs of the
 * pair distance, capped at upper_bound; candidates skipped when their AABB distance exceeds the running bound.
 * (The reference orders candidates by AABB distance and stops at the first candidate whose AABB distance exceeds the running
 * bound; once the bound is negative -- geometries inside each other's margins -- that makes its value order dependent.  The
 * oracle returns the order-independent quantity: the exact minimum over all enabled pairs.) */
double ko_distance(const ko_world* w, const double* q, double ub, int include_self, int32_t* pair, ko_counts* cnt) {
  int L=w->L; xf_t* T=(xf_t*)malloc(sizeof(xf_t)*L); fk_links(w,q,T);
  active_t* s1=(active_t*)malloc(sizeof(active_t)*(L+1));
  active_t* s2=(active_t*)malloc(sizeof(active_t)*(w->nterr+w->nobj+1));
  int n1=gather_links(w,T,s1), n2=gather_env(w,s2);
  compute_bbs(s1,n1,0.0); compute_bbs(s2,n2,0.0);
  double best=ub; if (pair) { pair[0]=-1; pair[1]=-1; }
  for (int i=0;i<n1;i++) for (int j=0;j<n2;j++) if (mask_en(w,s1[i].id,s2[j].id)||mask_en(w,s2[j].id,s1[i].id)) {
    if (cnt) cnt->n_box++;
    { double ad=aabb_dist(s1[i].lo,s1[i].hi,s2[j].lo,s2[j].hi); if (ad>0 && ad>=best) continue; }  /* touching boxes bound nothing: margins / radii can make the distance negative */
    double d=geom_pair_distance(s1[i].g,&s1[i].T,s2[j].g,&s2[j].T,best,cnt);
    if (d<best) { best=d; if (pair) { pair[0]=s1[i].id; pair[1]=s2[j].id; } } }
  if (include_self) for (int i=0;i<n1;i++) for (int j=i+1;j<n1;j++) if (mask_en(w,s1[i].id,s1[j].id)) {
    if (cnt) cnt->n_box++;
    { double ad=aabb_dist(s1[i].lo,s1[i].hi,s1[j].lo,s1[j].hi); if (ad>0 && ad>=best) continue; }
    double d=geom_pair_distance(s1[i].g,&s1[i].T,s1[j].g,&s1[j].T,best,cnt);
    if (d<best) { best=d; if (pair) { pair[0]=s1[i].id; pair[1]=s1[j].id; } } }
  free(T); free(s1); free(s2);
  return best;
}
void ko_distance_batch(const ko_world* w, const double* Q, int64_t N, double ub, int include_self, double* out_d, int32_t* out_pair, int nthreads) {
  int L=w->L; if (nthreads<=0) nthreads=ko_max_threads();
  #pragma omp parallel for schedule(dynamic,16) num_threads(nthreads)
  for (int64_t c=0;c<N;c++) { int32_t pr[2]; out_d[c]=ko_distance(w,Q+c*L,ub,include_self,pr,NULL);
    if (out_pair) { out_pair[2*c]=pr[0]; out_pair[2*c+1]=pr[1]; } }
}
