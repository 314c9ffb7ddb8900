/* TEST INFRASTRUCTURE -- CPU restatement of the reference's hot path in plain C.
 *
 * This file is the parity oracle where oracle/_ref (the reference's own code, compiled by
 * oracle/make_ref.py) is not available, and the "port" CPU baseline of bench.py.  It is NOT part
 * of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.
 *
 * PARITY PIN: tests/test_oracle_pin.py checks every function here against oracle/_ref built from
 * /root/reference/raygpu/kernel.cu (closest-hit ids and distances, float frames, parsed objects,
 * tree shape) on the reference's sample scenes, and against the golden vectors in tests/golden/
 * that were generated from oracle/_ref.  The reference itself ships no tests or golden vectors
 * (SURVEY.md section 4).
 *
 * Each function cites the kernel.cu lines it restates.  Arithmetic follows the C++ overloads the
 * host build of the reference selects (float pow/sqrt/tan for float arguments, double where a
 * double literal promotes the expression); compile with -O2 -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <dirent.h>
#include <pthread.h>

#include "philox_ref.h"

typedef struct { float x, y, z; } v3;

/* singleobject, kernel.cu:48-74 */
typedef struct {
    int type;
    v3 pos, rot, norm, n1, n2, n3, t1, t2, t3;
    int smooth, tex, mat;
    v3 dim, col;
    int texnum, rtexnum;
    v3 addional;
    int ncols;
} orc_object;

/* bvh, kernel.cu:79-96 (only the fields the device reads, plus the builder's bookkeeping) */
typedef struct {
    int children[2];
    int count, hit_node, miss_node, under;
    v3 min, max;
    int end, active;
} orc_node;

typedef struct { int w, h; unsigned char* rgba; } orc_tex;

typedef struct {
    orc_object* objs; int nobjs;          /* object lines */
    orc_node* nodes; int nnodes, cap_nodes;
    orc_tex* tex; int ntex; char** tex_paths;
    /* settings, kernel.cu:29-30, 119-132 */
    v3 campos, look; float aperture, focus, bg; int fov, max_depth, spp, backtex, W, H;
    uint64_t seed;
    uint64_t rays;
} orc_scene;

static v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static v3 V1(float a) { return V(a, a, a); }
static v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 mul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static v3 dvd(v3 a, v3 b) { return V(a.x / b.x, a.y / b.y, a.z / b.z); }
static float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                       /* :168-171 */
static v3 cross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); } /* :148-151 */
static float len(v3 a) { return sqrtf(dot(a, a)); }                                              /* :200-203 */
static v3 normalize(v3 v) { float inv = 1.0f / sqrtf(dot(v, v)); return V(v.x * inv, v.y * inv, v.z * inv); } /* :179-183 */

/* ---- intersection ------------------------------------------------------------------------ */
/* aabb2, kernel.cu:244-274 */
static int aabb2(v3 o, v3 d, v3 a, v3 b, float* dist)
{
    float t_min = 0, t_max = 10000;
    float origin[3] = { o.x, o.y, o.z }, direction[3] = { d.x, d.y, d.z }, mn[3] = { a.x, a.y, a.z }, mx[3] = { b.x, b.y, b.z };
    for (int k = 0; k < 3; k++) {
        float invD = 1.0f / direction[k];
        float t0 = (mn[k] - origin[k]) * invD, t1 = (mx[k] - origin[k]) * invD;
        if (invD < 0.0f) { float old = t0; t0 = t1; t1 = old; }
        t_min = t0 > t_min ? t0 : t_min;
        t_max = t1 < t_max ? t1 : t_max;
        if (t_max <= t_min) return 0;
    }
    *dist = t_min;
    return 1;
}

/* hit_tri, kernel.cu:277-313 */
static float hit_tri(v3 ro, v3 rd, v3 v0, v3 v1, v3 v2)
{
    const float EPSILON = 0.0001f;
    v3 edge1 = sub(v1, v0), edge2 = sub(v2, v0);
    v3 h = cross(rd, edge2);
    float a = dot(edge1, h);
    if (a > -EPSILON && a < EPSILON) return -1;
    float f = (float)(1.0 / (double)a);
    v3 s = sub(ro, v0);
    float u = f * dot(s, h);
    if (u < 0.0 || u > 1.0) return -1;
    v3 q = cross(s, edge1);
    float v = f * dot(rd, q);
    if (v < 0.0 || u + v > 1.0) return -1;
    float t = f * dot(edge2, q);
    return t > EPSILON ? t : -1;
}

/* hit_sphere, kernel.cu:316-333 */
static float hit_sphere(v3 center, float radius, v3 origin, v3 dir)
{
    v3 oc = sub(origin, center);
    float a = powf(len(dir), 2.0f);
    float half_b = dot(oc, dir);
    float c = powf(len(oc), 2.0f) - radius * radius;
    float disc = half_b * half_b - a * c;
    if (disc < 0) return -1.0f;
    return (-half_b - sqrtf(disc)) / a;
}

/* singlehit, kernel.cu:432-464: returns t or -1 */
static float singlehit(const orc_scene* s, v3 o, v3 d, int x)
{
    const orc_object* b = &s->objs[x];
    float dist;
    if (b->type == 0) dist = hit_sphere(b->pos, b->dim.x, o, d);
    else if (b->type == 2) dist = hit_tri(o, d, b->pos, b->dim, b->rot);
    else return -1;                                             /* the reference reads an uninitialised value here */
    if (dist < 10000 && dist > -0.0) return dist;
    return -1;
}

/* hit, kernel.cu:468-512: closest hit over the threaded tree */
static float hit(const orc_scene* s, v3 o, v3 d, int* id)
{
    float out = 10000000; int best = 0, found = 0;
    int box = s->nnodes > 0 ? 0 : -1;
    while (box != -1) {
        const orc_node* n = &s->nodes[box];
        float dister;
        int h = aabb2(o, d, n->min, n->max, &dister);
        if (h && dister < out) {
            if (n->end) {
                float t = singlehit(s, o, d, n->under);
                if (t > -0.01 && t < out) { out = t; best = n->under; found = 1; }
            }
            box = n->hit_node;
        } else box = n->miss_node;
    }
    if (!found) { *id = -1; return -1; }
    *id = best;
    return out;
}

/* ---- host tree build, kernel.cu:335-406, 1534-1909 ------------------------------------------ */
static void bounding_box(const orc_object* b, v3* mn, v3* mx)                                    /* :335-364 */
{
    if (b->type == 0) { *mn = sub(b->pos, V1(b->dim.x)); *mx = add(b->pos, V1(b->dim.x)); }
    else if (b->type == 2) {
        v3 a = b->pos, c = b->dim, e = b->rot;
        *mn = V((float)(fminf(a.x, fminf(c.x, e.x)) - 0.01), (float)(fminf(a.y, fminf(c.y, e.y)) - 0.01), (float)(fminf(a.z, fminf(c.z, e.z)) - 0.01));
        *mx = V((float)(fmaxf(a.x, fmaxf(c.x, e.x)) + 0.01), (float)(fmaxf(a.y, fmaxf(c.y, e.y)) + 0.01), (float)(fmaxf(a.z, fmaxf(c.z, e.z)) + 0.01));
    }
}
static void arraybound(const orc_scene* s, const int* idx, int n, v3* mn, v3* mx)                /* :383-406 */
{
    v3 tmn = V1(-1), tmx = V1(-1);
    for (int g = 0; g < n; g++) {
        bounding_box(&s->objs[idx[g]], &tmn, &tmx);
        if (g == 0) { *mn = tmn; *mx = tmx; }
        else {
            *mn = V(fminf(mn->x, tmn.x), fminf(mn->y, tmn.y), fminf(mn->z, tmn.z));
            *mx = V(fmaxf(mx->x, tmx.x), fmaxf(mx->y, tmx.y), fmaxf(mx->z, tmx.z));
        }
    }
}
typedef struct { float key; int idx; } orc_pair;
static int pair_cmp(const void* a, const void* b)                                                /* std::sort on pair<float,int>, :1547 */
{
    const orc_pair* p = (const orc_pair*)a; const orc_pair* q = (const orc_pair*)b;
    if (p->key < q->key) return -1;
    if (q->key < p->key) return 1;
    return (p->idx > q->idx) - (p->idx < q->idx);
}
static float sd_axis(const orc_object* objs, const int* idx, int n, int axis)                    /* calculateSD, :1560-1623 */
{
    float sum = 0.0f, sd = 0.0f;
    for (int i = 0; i < n; i++) sum += ((const float*)&objs[idx[i]].pos)[axis];
    float mean = sum / n;
    for (int i = 0; i < n; i++) sd = (float)(sd + pow((double)(((const float*)&objs[idx[i]].pos)[axis] - mean), 2.0));
    return sqrtf(sd / n);
}
/* split, :1678-1717 + sorto :1626-1674: sort `idx` in place along the axis of largest deviation of pos */
static void split_sort(const orc_scene* s, int* idx, int n)
{
    float dx = sd_axis(s->objs, idx, n, 0), dy = sd_axis(s->objs, idx, n, 1), dz = sd_axis(s->objs, idx, n, 2);
    float mx = fmaxf(dx, fmaxf(dy, dz));
    int axis = 0;
    if (mx == dx) axis = 0;
    if (mx == dy) axis = 1;
    if (mx == dz) axis = 2;
    orc_pair* p = (orc_pair*)malloc(sizeof(orc_pair) * (size_t)n);
    for (int i = 0; i < n; i++) { p[i].key = ((const float*)&s->objs[idx[i]].pos)[axis]; p[i].idx = idx[i]; }
    qsort(p, (size_t)n, sizeof(orc_pair), pair_cmp);
    for (int i = 0; i < n; i++) idx[i] = p[i].idx;
    free(p);
}
/* bvhr, :1745-1861: children are allocated consecutively, the left subtree is recursed first */
static void bvhr(orc_scene* s, int node, int* under, int count)
{
    if (s->nodes[node].end || count < 2) return;
    int p1 = count / 2, p2 = count - p1;
    split_sort(s, under, count);
    int an = s->nnodes++, bn = s->nnodes++;
    orc_node* A = &s->nodes[an]; orc_node* B = &s->nodes[bn];
    memset(A, 0, sizeof *A); memset(B, 0, sizeof *B);
    A->active = B->active = 1;
    A->count = p1; B->count = p2;
    if (p1 == 1) { A->under = under[0]; A->end = 1; }
    if (p2 == 1) { B->under = under[p1]; B->end = 1; }
    arraybound(s, under, p1, &A->min, &A->max);
    arraybound(s, under + p1, p2, &B->min, &B->max);
    s->nodes[node].children[0] = an; s->nodes[node].children[1] = bn;
    bvhr(s, an, under, p1);
    bvhr(s, bn, under + p1, p2);
}
static void build_links(orc_scene* s, int self, int next_right)                                   /* :1720-1742 */
{
    orc_node* n = &s->nodes[self];
    if (!n->end) {
        n->hit_node = n->children[0]; n->miss_node = next_right;
        build_links(s, n->children[0], n->children[1]);
        build_links(s, n->children[1], next_right);
    } else { n->hit_node = next_right; n->miss_node = next_right; }
}
/* build_bvh, :1864-1909.  Objects of a type the reference cannot bound are left out (it would reuse the
 * previous object's box); the phantom object at index `lines` is not modelled. */
static void build_bvh(orc_scene* s)
{
    int* under = (int*)malloc(sizeof(int) * (size_t)(s->nobjs + 1));
    int n = 0;
    for (int i = 0; i < s->nobjs; i++) {
        const orc_object* o = &s->objs[i];
        int ok = (o->type == 2 && o->ncols >= 16) || (o->type == 0 && o->ncols >= 10);
        if (ok) under[n++] = i;
    }
    free(s->nodes);
    s->cap_nodes = 2 * n + 2;
    s->nodes = (orc_node*)calloc((size_t)s->cap_nodes, sizeof(orc_node));
    s->nnodes = 0;
    if (n > 0) {
        s->nnodes = 1;
        orc_node* r = &s->nodes[0];
        r->active = 1; r->count = n; r->end = 0;
        arraybound(s, under, n, &r->min, &r->max);
        if (n == 1) { r->end = 1; r->under = under[0]; }
        bvhr(s, 0, under, n);
        build_links(s, 0, -1);
    }
    free(under);
}

/* ---- textures --------------------------------------------------------------------------------- */
/* tex2D<uchar4> with the descriptor of kernel.cu:1959-1964: normalised, wrap, point */
static void tex2d(const orc_scene* s, int k, float u, float v, unsigned char out[4])
{
    out[0] = out[1] = out[2] = out[3] = 0;
    if (k < 0 || k >= s->ntex) return;
    const orc_tex* t = &s->tex[k];
    if (t->w <= 0 || t->h <= 0 || !t->rgba) return;
    float fu = u - floorf(u), fv = v - floorf(v);
    int ix = (int)floorf(fu * (float)t->w), iy = (int)floorf(fv * (float)t->h);
    if (!(ix >= 0)) ix = 0;
    if (!(iy >= 0)) iy = 0;
    if (ix >= t->w) ix = t->w - 1;
    if (iy >= t->h) iy = t->h - 1;
    memcpy(out, t->rgba + 4 * ((size_t)iy * t->w + ix), 4);
}

/* ---- sampling, kernel.cu:640-662, 988-994 ------------------------------------------------------ */
/* The reference draws all components inside one make_float3(...) argument list (unspecified
 * order); the host build of the reference evaluates it right to left, restated explicitly here. */
static v3 random_in_unit_sphere(orc_rng* r)
{
    for (;;) {
        float c = (float)((double)orc_rng_uniform(r) * 2 - 1);
        float b = (float)((double)orc_rng_uniform(r) * 2 - 1);
        float a = (float)((double)orc_rng_uniform(r) * 2 - 1);
        v3 p = V(a, b, c);
        if (powf(len(p), 2.0f) >= 1) continue;
        return p;
    }
}
static v3 random_in_unit_disk(orc_rng* r)
{
    for (;;) {
        float b = (float)(((double)orc_rng_uniform(r) * 2) - 1);
        float a = (float)(((double)orc_rng_uniform(r) * 2) - 1);
        v3 p = V(a, b, 0);
        if (powf(len(p), 2.0f) >= 1) continue;
        return p;
    }
}

/* ---- shading ------------------------------------------------------------------------------------ */
static v3 reflect(v3 v, v3 n) { float k = (float)(2.0 * (double)dot(v, n)); return sub(v, mul(V1(k), n)); }          /* :667-669 */
static v3 refract(v3 uv, v3 n, float eta)                                                                          /* :678-683 */
{
    float cos_theta = (float)fmin((double)dot(mul(uv, V1(-1)), n), 1.0);
    v3 perp = mul(V1(eta), add(uv, mul(V1(cos_theta), n)));
    float k = (float)(-sqrt(fabs(1.0 - (double)powf(len(perp), 2.0f))));
    return add(perp, mul(V1(k), n));
}
static float reflectance(float cosine, float ref_idx)                                                              /* :686-691 */
{
    float r0 = (1 - ref_idx) / (1 + ref_idx);
    r0 = r0 * r0;
    return (float)((double)r0 + (double)(1 - r0) * pow((double)(1 - cosine), 5.0));
}
/* getnormal, :703-773 */
static v3 getnormal(const orc_scene* s, int obj, v3 origin, v3 hitpoint, v3 dir, v3* texco)
{
    const orc_object* b = &s->objs[obj];
    if (b->type == 0) return dvd(sub(hitpoint, b->pos), V1(b->dim.x));
    if (b->type == 2) {
        v3 v0 = b->pos, v1 = b->dim, v2 = b->rot;
        v3 e1 = sub(v1, v0), e2 = sub(v2, v0);
        v3 N = cross(e1, e2);
        v3 pvec = cross(dir, e2);
        float det = dot(e1, pvec);
        float inv = 1 / det;
        v3 tvec = sub(origin, v0);
        float ux = dot(tvec, pvec) * inv;
        v3 qvec = cross(tvec, e1);
        float uy = dot(dir, qvec) * inv;
        float uz = 1 - ux - uy;
        *texco = add(add(mul(V1(uz), b->t1), mul(V1(ux), b->t2)), mul(V1(uy), b->t3));
        if (b->norm.z != -20) {
            N = b->norm;
            if (b->n1.z != -20 && b->smooth) N = add(add(mul(V1(uz), b->n1), mul(V1(ux), b->n2)), mul(V1(uy), b->n3));
        }
        return normalize(N);
    }
    return normalize(sub(hitpoint, b->pos));
}
/* raycolor, :787-982 */
static v3 raycolor(orc_scene* s, v3 origin, v3 dir, int max_depth, orc_rng* rng, uint64_t* rays)
{
    v3 raydir = dir, rayo = origin, att = V1(1.0f);
    for (int i = 0; i < max_depth; i++) {
        v3 texco = V1(0);
        int g;
        (*rays)++;
        float t = hit(s, rayo, raydir, &g);
        if (t > 0.0) {
            v3 hitpoint = add(rayo, mul(V1(t), raydir));
            v3 N = getnormal(s, g, rayo, hitpoint, raydir, &texco);
            int inorout = dot(raydir, N) < 0;
            if (!inorout) N = mul(N, V1(-1));
            const orc_object* b = &s->objs[g];
            v3 ocolor = b->col;
            float rough = b->addional.y;
            unsigned char C[4];
            if (b->texnum >= 0) { tex2d(s, b->texnum, texco.x, -texco.y + 1, C); ocolor = V((float)C[0] / 255, (float)C[1] / 255, (float)C[2] / 255); }
            else if (b->tex) {                                                                          /* checker, :776-784 */
                float yes = floorf(texco.x * 10) + floorf(texco.y * 10);
                ocolor = fmodf(yes, 2.0f) == 0 ? V1(0.8f) : b->col;
            }
            if (b->rtexnum >= 0) { tex2d(s, b->rtexnum, texco.x, -texco.y + 1, C); rough = (float)C[0] / 255 / 2; }
            if (b->mat == 0) {
                v3 target = add(hitpoint, N);
                if (b->addional.x == 0) target = add(target, random_in_unit_sphere(rng));
                else target = add(target, normalize(random_in_unit_sphere(rng)));
                att = mul(att, ocolor); rayo = hitpoint; raydir = normalize(sub(target, hitpoint));
            } else if (b->mat == 2) {
                att = mul(att, ocolor); rayo = hitpoint; raydir = reflect(normalize(raydir), N);
            } else if (b->mat == 3) {
                v3 refl = reflect(normalize(raydir), N);
                att = mul(att, ocolor); rayo = hitpoint;
                raydir = add(refl, mul(V1(rough), random_in_unit_sphere(rng)));
            } else if (b->mat == 5) {
                float pick = orc_rng_uniform(rng);
                if (pick > 0.8) {
                    v3 refl = reflect(normalize(raydir), N);
                    att = mul(att, ocolor); rayo = hitpoint;
                    raydir = add(refl, mul(V1(rough), random_in_unit_sphere(rng)));
                } else {
                    v3 target = add(add(hitpoint, N), random_in_unit_sphere(rng));
                    att = mul(att, ocolor); rayo = hitpoint; raydir = normalize(sub(target, hitpoint));
                }
            } else if (b->mat == 4) {
                float ir = b->addional.y;
                float ratio = inorout ? (float)(1.0 / (double)ir) : ir;
                v3 unit = normalize(raydir);
                float cos_theta = (float)fmin((double)dot(mul(unit, V1(-1)), N), 1.0);
                float sin_theta = (float)sqrt(1.0 - (double)(cos_theta * cos_theta));
                int cannot = (ratio * sin_theta) > 1.0;
                v3 nd;
                if (cannot || reflectance(cos_theta, ratio) > orc_rng_uniform(rng)) nd = reflect(unit, N);
                else nd = refract(unit, N, ratio);
                att = mul(att, ocolor); rayo = hitpoint; raydir = nd;
            } else {
                return mul(ocolor, att);
            }
        } else {
            v3 ud = normalize(raydir);
            if (s->backtex > -1) {
                double dx = ud.x, dy = ud.y, dz = (double)ud.z + 1.;
                float m = (float)(2. * sqrt(dx * dx + dy * dy + dz * dz));
                v3 tc = add(dvd(ud, V1(m)), V1(.5f));
                tc.y = -tc.y;
                unsigned char C[4];
                tex2d(s, s->backtex, tc.x, -tc.y + 1, C);
                v3 c2 = V((float)C[0] / 255, (float)C[1] / 255, (float)C[2] / 255);
                return mul(mul(att, c2), V1(s->bg));
            }
            float tt = (float)(0.5 * ((double)ud.y + 1.0));
            float omt = (float)(1.0 - (double)tt);
            v3 c = add(mul(V1(omt), V1(1.0f)), mul(V1(tt), V(0.5f, 0.7f, 1.0f)));
            return mul(mul(att, c), V1(s->bg));
        }
    }
    return V1(0.0f);
}

/* ---- camera + per-pixel loop: Kernel, :998-1093 -------------------------------------------------- */
typedef struct { v3 from, llc, hor, ver, uu, vu; float lens, wdiv, hdiv; } orc_camera;
static orc_camera make_camera(const orc_scene* s, int divisor)
{
    orc_camera c;
    float div = (float)divisor;
    c.wdiv = (float)(s->W / div); c.hdiv = (float)(s->H / div);
    float aspect = c.wdiv / c.hdiv;
    float fov = (float)((float)s->fov * M_PI / 180);
    float vh = (float)(2.0 * tanf(fov / 2));
    float vw = aspect * vh;
    v3 wu = normalize(sub(s->campos, s->look));
    v3 uu = normalize(cross(V(0, 1, 0), wu));
    v3 vu = cross(wu, uu);
    c.hor = mul(mul(V1(s->focus), V1(vw)), uu);
    c.ver = mul(mul(V1(s->focus), V1(vh)), vu);
    c.llc = sub(sub(sub(s->campos, dvd(c.hor, V1(2))), dvd(c.ver, V1(2))), mul(V1(s->focus), wu));
    c.from = s->campos; c.uu = uu; c.vu = vu;
    c.lens = s->aperture / 2;
    return c;
}
static void camera_ray(const orc_camera* c, int x, int y, orc_rng* rng, v3* o, v3* d)
{
    float nu = (float)(((double)(float)x + (double)orc_rng_uniform(rng)) / (double)c->wdiv);
    float nv = (float)(((double)(float)y + (double)orc_rng_uniform(rng)) / (double)c->hdiv);
    v3 rd = mul(V1(c->lens), random_in_unit_disk(rng));
    v3 off = add(mul(c->uu, V1(rd.x)), mul(c->vu, V1(rd.y)));
    *d = sub(sub(add(add(c->llc, mul(V1(nu), c->hor)), mul(V1(nv), c->ver)), c->from), off);
    *o = add(c->from, off);
}

typedef struct { orc_scene* s; float* out_f; int* out_i; int divisor; unsigned sample_base; int t, nt; uint64_t rays; } frame_job;
static void* frame_worker(void* arg)
{
    frame_job* j = (frame_job*)arg;
    orc_scene* s = j->s;
    orc_camera cam = make_camera(s, j->divisor);
    const int gw = s->W / j->divisor / 8 * 8, gh = s->H / j->divisor / 8 * 8;     /* launched grid, :2634-2636 */
    const float scale = (float)(1.0 / (double)(float)s->spp);
    for (int x = j->t; x < gw; x += j->nt)
        for (int y = 0; y < gh; y++) {
            v3 sum = V1(0);
            for (int k = 0; k < s->spp; k++) {
                orc_rng rng;
                orc_rng_init(&rng, s->seed, (uint32_t)x, (uint32_t)y, j->sample_base + (uint32_t)k);
                v3 o, d;
                camera_ray(&cam, x, y, &rng, &o, &d);
                sum = add(sum, raycolor(s, o, d, s->max_depth, &rng, &j->rays));
            }
            size_t w = ((size_t)x * s->H + y) * 3;
            float v[3] = { sum.x * 255 * scale, sum.y * 255 * scale, sum.z * 255 * scale };
            for (int c = 0; c < 3; c++) {
                if (j->out_f) j->out_f[w + c] = v[c];
                if (j->out_i) j->out_i[w + c] = (v[c] != v[c]) ? 0 : (v[c] >= 2147483648.0f ? INT32_MAX : (v[c] <= -2147483648.0f ? INT32_MIN : (int)v[c]));
            }
        }
    return NULL;
}

/* ---- .rts / .ppm ingest: getnum + read + gettexnum, kernel.cu:1113-1530, 1172-1183 ------------------ */
static int find_texture(const orc_scene* s, const char* q)
{
    for (int i = 0; i < s->ntex; i++) {
        char low[4096]; size_t n = strlen(s->tex_paths[i]);
        if (n >= sizeof low) n = sizeof low - 1;
        for (size_t k = 0; k < n; k++) low[k] = (char)tolower((unsigned char)s->tex_paths[i][k]);
        low[n] = 0;
        if (strstr(low, q)) return i;
    }
    return -1;
}
static int load_ppm(const char* path, orc_tex* t)
{
    FILE* f = fopen(path, "rb");
    if (!f) return 0;
    int c0 = fgetc(f), c1 = fgetc(f), vals[3], got = 0;
    if (c0 != 'P' || (c1 != '6' && c1 != '5')) { fclose(f); return 0; }
    int ch = c1 == '6' ? 3 : 1;
    while (got < 3) {
        int c = fgetc(f);
        if (c == EOF) { fclose(f); return 0; }
        if (c == '#') { while (c != '\n' && c != EOF) c = fgetc(f); continue; }
        if (isspace(c)) continue;
        ungetc(c, f);
        if (fscanf(f, "%d", &vals[got]) != 1) { fclose(f); return 0; }
        got++;
    }
    fgetc(f);
    size_t n = (size_t)vals[0] * vals[1];
    unsigned char* raw = (unsigned char*)malloc(n * ch);
    if (fread(raw, 1, n * ch, f) != n * ch) { free(raw); fclose(f); return 0; }
    fclose(f);
    t->w = vals[0]; t->h = vals[1];
    t->rgba = (unsigned char*)calloc(n, 4);
    for (size_t i = 0; i < n; i++) for (int k = 0; k < 3; k++) t->rgba[4 * i + k] = raw[ch == 3 ? 3 * i + k : i];
    free(raw);
    return 1;
}
static int cmp_str(const void* a, const void* b) { return strcmp(*(char* const*)a, *(char* const*)b); }
static void scan_textures(orc_scene* s, const char* dir)
{
    if (!dir || !dir[0]) return;
    DIR* d = opendir(dir);
    if (!d) return;
    struct dirent* e;
    int cap = 0;
    while ((e = readdir(d))) {
        if (!strstr(e->d_name, "ppm") && !strstr(e->d_name, "PPM")) continue;
        if (s->ntex == cap) { cap = cap ? 2 * cap : 16; s->tex_paths = (char**)realloc(s->tex_paths, sizeof(char*) * (size_t)cap); }
        size_t n = strlen(dir) + strlen(e->d_name) + 2;
        s->tex_paths[s->ntex] = (char*)malloc(n);
        snprintf(s->tex_paths[s->ntex], n, "%s/%s", dir, e->d_name);
        s->ntex++;
    }
    closedir(d);
    qsort(s->tex_paths, (size_t)s->ntex, sizeof(char*), cmp_str);
    s->tex = (orc_tex*)calloc((size_t)(s->ntex ? s->ntex : 1), sizeof(orc_tex));
    for (int i = 0; i < s->ntex; i++) load_ppm(s->tex_paths[i], &s->tex[i]);
}
static void object_defaults(orc_object* o)                                                        /* :48-74 */
{
    memset(o, 0, sizeof *o);
    o->norm = o->n1 = o->n2 = o->n3 = V(-2, -3, -20);
    o->t1 = V(0, 1, 0); o->t2 = V(0, 0, 0); o->t3 = V(1, 0, 0);
    o->texnum = o->rtexnum = -1;
}
static void set_object_col(orc_scene* s, orc_object* o, int col, const char* tok)                 /* :1316-1503 */
{
    float f = strtof(tok, NULL); int iv = (int)strtol(tok, NULL, 10);
    switch (col) {
    case 0: o->pos.x = f; break; case 1: o->pos.y = f; break; case 2: o->pos.z = f; break;
    case 3: o->type = iv; break;
    case 4: o->col.x = f; break; case 5: o->col.y = f; break; case 6: o->col.z = f; break;
    case 7: o->addional.y = f; break; case 8: o->addional.x = f; break;
    case 9: o->dim.x = f; break; case 10: o->dim.y = f; break; case 11: o->dim.z = f; break;
    case 12: o->mat = iv; break;
    case 13: o->rot.x = f; break; case 14: o->rot.y = f; break; case 15: o->rot.z = f; break;
    case 16: o->norm.x = f; break; case 17: o->norm.y = f; break; case 18: o->norm.z = f; break;
    case 19: o->n1.x = f; break; case 20: o->n1.y = f; break; case 21: o->n1.z = f; break;
    case 22: o->n2.x = f; break; case 23: o->n2.y = f; break; case 24: o->n2.z = f; break;
    case 25: o->n3.x = f; break; case 26: o->n3.y = f; break; case 27: o->n3.z = f; break;
    case 28: o->t1.x = f; break; case 29: o->t1.y = f; break;
    case 30: o->t2.x = f; break; case 31: o->t2.y = f; break;
    case 32: o->t3.x = f; break; case 33: o->t3.y = f; break;
    case 34: o->smooth = iv == 1; break;
    case 35: o->tex = iv == 1; break;
    case 36: if (strcmp(tok, "no") != 0) o->texnum = find_texture(s, tok); break;
    case 37: if (strcmp(tok, "no") != 0) o->rtexnum = find_texture(s, tok); break;
    default: break;
    }
}
static void set_settings_col(orc_scene* s, int col, const char* tok)                              /* :1230-1293 */
{
    float f = strtof(tok, NULL); int iv = (int)strtol(tok, NULL, 10);
    switch (col) {
    case 1: s->campos.x = f; break; case 2: s->campos.y = f; break; case 3: s->campos.z = f; break;
    case 4: s->aperture = f; break;
    case 5: s->look.x = f; break; case 6: s->look.y = f; break; case 7: s->look.z = f; break;
    case 8: s->focus = f; break; case 9: s->fov = iv; break; case 10: s->max_depth = iv; break; case 11: s->spp = iv; break;
    case 12: s->bg = f; break;
    case 13: if (strcmp(tok, "no") != 0) s->backtex = find_texture(s, tok); break;
    case 14: s->W = iv; break; case 15: s->H = iv; break;
    default: break;
    }
}

/* ---- C entry points (loaded with ctypes by tests/ and bench.py) ------------------------------------- */
void orc_free(orc_scene* s)
{
    if (!s) return;
    free(s->objs); free(s->nodes);
    for (int i = 0; i < s->ntex; i++) { free(s->tex[i].rgba); free(s->tex_paths[i]); }
    free(s->tex); free(s->tex_paths);
    free(s);
}

orc_scene* orc_load(const char* rts_path, const char* tex_dir)
{
    FILE* f = fopen(rts_path, "rb");
    if (!f) return NULL;
    orc_scene* s = (orc_scene*)calloc(1, sizeof *s);
    s->campos = V(0, 0, 2); s->look = V(0, 0, 0); s->aperture = 0.01f; s->focus = 3; s->bg = 1;
    s->fov = 45; s->max_depth = 50; s->spp = 1; s->backtex = -1; s->W = 1280; s->H = 720;
    scan_textures(s, tex_dir);
    size_t cap = 0; char* line = NULL; size_t lcap = 0; ssize_t n;
    while ((n = getline(&line, &lcap, f)) >= 0) {
        while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
        if (n == 0 || line[0] == '/') continue;
        int is_settings = line[0] == '*';
        orc_object* o = NULL;
        if (!is_settings) {
            if ((size_t)s->nobjs == cap) { cap = cap ? cap * 2 : 1024; s->objs = (orc_object*)realloc(s->objs, cap * sizeof(orc_object)); }
            o = &s->objs[s->nobjs++];
            object_defaults(o);
        }
        int col = 0; char* p = line;
        for (;;) {
            char* c = strchr(p, ',');
            if (c) *c = 0;
            if (is_settings) set_settings_col(s, col, p); else set_object_col(s, o, col, p);
            col++;
            if (!c) break;
            p = c + 1;
        }
        if (o) o->ncols = col;
    }
    free(line); fclose(f);
    build_bvh(s);
    return s;
}

int orc_num_objects(const orc_scene* s) { return s->nobjs; }
int orc_num_nodes(const orc_scene* s) { return s->nnodes; }
int orc_sizeof_object(void) { return (int)sizeof(orc_object); }
const orc_object* orc_objects(const orc_scene* s) { return s->objs; }
/* out[16]: cam xyz, aperture, look xyz, focus, fov, depth, spp, bg, backtex, W, H, 0 */
void orc_get_settings(const orc_scene* s, float* o)
{
    o[0] = s->campos.x; o[1] = s->campos.y; o[2] = s->campos.z; o[3] = s->aperture; o[4] = s->look.x; o[5] = s->look.y; o[6] = s->look.z;
    o[7] = s->focus; o[8] = (float)s->fov; o[9] = (float)s->max_depth; o[10] = (float)s->spp; o[11] = s->bg; o[12] = (float)s->backtex;
    o[13] = (float)s->W; o[14] = (float)s->H; o[15] = 0;
}
void orc_set_settings(orc_scene* s, const float* in)
{
    s->campos = V(in[0], in[1], in[2]); s->aperture = in[3]; s->look = V(in[4], in[5], in[6]); s->focus = in[7];
    s->fov = (int)in[8]; s->max_depth = (int)in[9]; s->spp = (int)in[10]; s->bg = in[11]; s->backtex = (int)in[12];
    s->W = (int)in[13]; s->H = (int)in[14];
}
void orc_set_seed(orc_scene* s, uint64_t seed) { s->seed = seed; }

void orc_hit(const orc_scene* s, const float* o3, const float* d3, int n, float* t, int* id)
{
    for (int i = 0; i < n; i++) {
        int g; float tt = hit(s, V(o3[3 * i], o3[3 * i + 1], o3[3 * i + 2]), V(d3[3 * i], d3[3 * i + 1], d3[3 * i + 2]), &g);
        int ok = tt > 0.0;
        t[i] = tt; id[i] = ok ? g : -1;
    }
}
/* brute force over every object: the BVH-independent definition of the closest hit */
void orc_hit_brute(const orc_scene* s, const float* o3, const float* d3, int n, float* t, int* id)
{
    for (int i = 0; i < n; i++) {
        v3 o = V(o3[3 * i], o3[3 * i + 1], o3[3 * i + 2]), d = V(d3[3 * i], d3[3 * i + 1], d3[3 * i + 2]);
        float best = 10000000; int g = -1;
        for (int k = 0; k < s->nobjs; k++) {
            float tt = singlehit(s, o, d, k);
            if (tt > -0.01 && tt < best) { best = tt; g = k; }
        }
        int ok = g >= 0 && best > 0.0;
        t[i] = g >= 0 ? best : -1; id[i] = ok ? g : -1;
    }
}

/* Closest hit by definition over TRIANGLE ARRAYS (no scene file, no tree): hit_tri (kernel.cu:277-313) against every
 * triangle, acceptance as singlehit / hit (:449, :488: 0 < t < 10000, strictly closer wins, so the lowest index keeps
 * an exact tie).  For scenes too large to push through the text format in a test (10 M triangles): rays are split
 * over `threads` host threads.  v0/v1/v2: n*3 floats; t: nrays (or -1), id: nrays (or -1). */
typedef struct { const float *v0, *v1, *v2, *o3, *d3; int n, r0, r1; float* t; int* id; } brute_job;
static void* brute_worker(void* arg)
{
    brute_job* j = (brute_job*)arg;
    for (int i = j->r0; i < j->r1; i++) {
        v3 o = V(j->o3[3 * i], j->o3[3 * i + 1], j->o3[3 * i + 2]), d = V(j->d3[3 * i], j->d3[3 * i + 1], j->d3[3 * i + 2]);
        float best = 10000; int g = -1;
        for (int k = 0; k < j->n; k++) {
            const float *a = j->v0 + 3 * (size_t)k, *b = j->v1 + 3 * (size_t)k, *c = j->v2 + 3 * (size_t)k;
            float tt = hit_tri(o, d, V(a[0], a[1], a[2]), V(b[0], b[1], b[2]), V(c[0], c[1], c[2]));
            if (tt > -0.0 && tt < best) { best = tt; g = k; }
        }
        j->t[i] = g >= 0 ? best : -1; j->id[i] = g;
    }
    return NULL;
}
void orc_brute_tris(const float* v0, const float* v1, const float* v2, int n, const float* o3, const float* d3, int nrays, float* t, int* id, int threads)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (threads > nrays) threads = nrays > 0 ? nrays : 1;
    pthread_t th[256]; brute_job jobs[256];
    for (int k = 0; k < threads; k++) {
        brute_job j = { v0, v1, v2, o3, d3, n, (int)((long long)nrays * k / threads), (int)((long long)nrays * (k + 1) / threads), t, id };
        jobs[k] = j;
        if (k > 0) pthread_create(&th[k], NULL, brute_worker, &jobs[k]);
    }
    brute_worker(&jobs[0]);
    for (int k = 1; k < threads; k++) pthread_join(th[k], NULL);
}

/* one frame; out_f/out_i indexed (x*H + y)*3 like outputr; returns rays traced */
uint64_t orc_frame(orc_scene* s, float* out_f, int* out_i, int divisor, unsigned sample_base, int threads)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256]; frame_job jobs[256];
    for (int t = 0; t < threads; t++) {
        frame_job j = { s, out_f, out_i, divisor, sample_base, t, threads, 0 };
        jobs[t] = j;
        if (t > 0) pthread_create(&th[t], NULL, frame_worker, &jobs[t]);
    }
    frame_worker(&jobs[0]);
    uint64_t rays = jobs[0].rays;
    for (int t = 1; t < threads; t++) { pthread_join(th[t], NULL); rays += jobs[t].rays; }
    s->rays = rays;
    return rays;
}

void orc_primary_rays(const orc_scene* s, unsigned sample, float* o3, float* d3)
{
    orc_camera cam = make_camera(s, 1);
    for (int y = 0; y < s->H; y++)
        for (int x = 0; x < s->W; x++) {
            orc_rng rng; orc_rng_init(&rng, s->seed, (uint32_t)x, (uint32_t)y, sample);
            v3 o, d; camera_ray(&cam, x, y, &rng, &o, &d);
            size_t k = ((size_t)y * s->W + x) * 3;
            o3[k] = o.x; o3[k + 1] = o.y; o3[k + 2] = o.z; d3[k] = d.x; d3[k + 1] = d.y; d3[k + 2] = d.z;
        }
}

uint32_t orc_philox_word(uint64_t seed, uint32_t x, uint32_t y, uint32_t sample, uint32_t n)
{
    uint32_t ctr[4] = { x, y, sample, n >> 2 }, out[4];
    orc_philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), ctr, out);
    return out[n & 3u];
}
