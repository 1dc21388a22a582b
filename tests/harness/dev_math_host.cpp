/* Host build of tmkit_b200/csrc/dev_math.cuh for CPU-side tests (TEST INFRASTRUCTURE: lets the
 * FP32 certificates and the FP64 exact path be checked against the oracle without a GPU; it is
 * never part of libtmkcl.so).  Dimensions use the DEVICE convention: sphere r | box half
 * extents | cylinder (radius, half height), centred. */
#include "dev_math.cuh"
using namespace tmk;

template <typename T> static Xf<T> xf_of(const double *p) {
    Xf<T> x;
    x.r.x = (T)p[0]; x.r.y = (T)p[1]; x.r.z = (T)p[2]; x.r.w = (T)p[3];
    x.v.x = (T)p[4]; x.v.y = (T)p[5]; x.v.z = (T)p[6];
    return x;
}
template <typename T> static Shape<T> shape_of(int k, const double *d) {
    Shape<T> s;
    s.kind = k; s.a = (T)d[0]; s.b = (T)d[1]; s.c = (T)d[2];
    return s;
}
template <typename T> static int convex(int sa, const double *da, const double *ta, int sb, const double *db, const double *tb) {
    return test_convex(shape_of<T>(sa, da), xf_of<T>(ta), shape_of<T>(sb, db), xf_of<T>(tb));
}
template <typename T> static int tri_shape(int s, const double *d, const double *tri) {
    return test_tri_vs_shape(shape_of<T>(s, d), mk<T>((T)tri[0], (T)tri[1], (T)tri[2]), mk<T>((T)tri[3], (T)tri[4], (T)tri[5]),
                             mk<T>((T)tri[6], (T)tri[7], (T)tri[8]));
}
extern "C" {
int h_convex_f32(int sa, const double *da, const double *ta, int sb, const double *db, const double *tb) { return convex<float>(sa, da, ta, sb, db, tb); }
int h_convex_f64(int sa, const double *da, const double *ta, int sb, const double *db, const double *tb) { return convex<double>(sa, da, ta, sb, db, tb); }
int h_tri_shape_f32(int s, const double *d, const double *tri) { return tri_shape<float>(s, d, tri); }
int h_tri_shape_f64(int s, const double *d, const double *tri) { return tri_shape<double>(s, d, tri); }
int h_tri_tri_f64(const double *a, const double *b) {
    return test_tri_tri(mk<double>(a[0], a[1], a[2]), mk<double>(a[3], a[4], a[5]), mk<double>(a[6], a[7], a[8]),
                        mk<double>(b[0], b[1], b[2]), mk<double>(b[3], b[4], b[5]), mk<double>(b[6], b[7], b[8]));
}
/* quick certificate of the FP32 broadphase stage: dims in the device convention, brad computed here */
int h_quick_f32(int sa, const double *da, const double *ta, int sb, const double *db, const double *tb) {
    auto brad = [](int k, const double *d) {
        if (k == K_SPHERE) return (float)d[0];
        if (k == K_BOX) return (float)(sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) * (1 + 1e-6));
        return (float)(sqrt(d[0] * d[0] + d[1] * d[1]) * (1 + 1e-6));
    };
    QG A = qg_make(sa, (float)da[0], (float)da[1], (float)da[2], brad(sa, da), xf_of<float>(ta), mk<float>(0.f, 0.f, 0.f));
    QG B = qg_make(sb, (float)db[0], (float)db[1], (float)db[2], brad(sb, db), xf_of<float>(tb), mk<float>(0.f, 0.f, 0.f));
    return quick_pair(sa, A, sb, B);
}
int h_convex_f32_iters(int sa, const double *da, const double *ta, int sb, const double *db, const double *tb, int *iters) {
    *iters = 0;
    return test_convex(shape_of<float>(sa, da), xf_of<float>(ta), shape_of<float>(sb, db), xf_of<float>(tb), iters);
}
int h_convex_f64_iters(int sa, const double *da, const double *ta, int sb, const double *db, const double *tb, int *iters) {
    *iters = 0;
    return test_convex(shape_of<double>(sa, da), xf_of<double>(ta), shape_of<double>(sb, db), xf_of<double>(tb), iters);
}
int h_quick_tri_f32(int s, const double *d, const double *tri) {
    return quick_tri(shape_of<float>(s, d), mk<float>((float)tri[0], (float)tri[1], (float)tri[2]),
                     mk<float>((float)tri[3], (float)tri[4], (float)tri[5]), mk<float>((float)tri[6], (float)tri[7], (float)tri[8]));
}
int h_quick_tri_f64(int s, const double *d, const double *tri) {
    return quick_tri(shape_of<double>(s, d), mk<double>(tri[0], tri[1], tri[2]), mk<double>(tri[3], tri[4], tri[5]),
                     mk<double>(tri[6], tri[7], tri[8]));
}
int h_ompl_mid(int nd, int k) { return ompl_mid(nd, k); }
/* FK chain in float vs double: returns the absolute pose of a serial revolute chain */
void h_fk_chain(int n, const double *E, const double *axis, const double *q, int use_float, double *out) {
    if (use_float) {
        Xf<float> cur; bool have = false;
        for (int i = 0; i < n; i++) {
            Xf<float> e = xf_of<float>(E + 7 * i);
            Xf<float> r = fk_joint(e, mk<float>((float)axis[3*i], (float)axis[3*i+1], (float)axis[3*i+2]), 0.f, 1, (float)q[i], have ? &cur : (const Xf<float>*)nullptr);
            cur = r; have = true;
        }
        out[0]=cur.r.x; out[1]=cur.r.y; out[2]=cur.r.z; out[3]=cur.r.w; out[4]=cur.v.x; out[5]=cur.v.y; out[6]=cur.v.z;
    } else {
        Xf<double> cur; bool have = false;
        for (int i = 0; i < n; i++) {
            Xf<double> e = xf_of<double>(E + 7 * i);
            Xf<double> r = fk_joint(e, mk<double>(axis[3*i], axis[3*i+1], axis[3*i+2]), 0.0, 1, q[i], have ? &cur : (const Xf<double>*)nullptr);
            cur = r; have = true;
        }
        out[0]=cur.r.x; out[1]=cur.r.y; out[2]=cur.r.z; out[3]=cur.r.w; out[4]=cur.v.x; out[5]=cur.v.y; out[6]=cur.v.z;
    }
}
}
