/* TEST INFRASTRUCTURE ONLY -- part of oracle/ (see oracle/README.md).  Never linked into the product.
 *
 * Plain-C restatement of the reference's neighbour-query + pair-accumulation arithmetic (freud @
 * e4272dbe).  It exists so that (a) the checker does not depend on /root/reference being present,
 * (b) the LinkCell ("wrap") flavour can be checked at N = 1e6, where the reference's own LinkCell is
 * quadratic, and (c) the CPU baseline has a sane cell loop to time.
 *
 * PARITY PINNING: tests/test_oracle_port.py checks every function below bit for bit against
 * oracle/_ref/libfreud_ref.so (the unmodified reference compiled here) on cubic, orthorhombic,
 * triclinic and 2-D boxes, and against the reference tests' known answers.
 *
 * Every float32 operation is written as its own statement so that no contraction can occur; the file
 * is compiled with -ffp-contract=off and without -march (oracle/Makefile).
 *
 * Equivalences used (verified against the compiled reference):
 *   E1  LinkCell ball query   == { (i,j) : r = Box::wrap(p_j - q_i), r_min^2 <= r.r < r_max^2 }
 *   E2  AABBQuery ball query  == { (i,j,k) : r = p_j - (q_i + image_k), same window }, 27 (9) images
 *   E3  AABBQuery kNN         == k smallest closest-image distances in the E2 arithmetic
 *   E5  CellQuery ball query  == { (i,j,w) : r = (p_j + shift_w) - q_i, same window }, shift_w the ghost
 *       displacement of CellQuery.h:246-262 (points and queries inside the box, r_max <= the r_max the grid
 *       was built for)
 * so candidate generation (the cell grid below) is free to differ from the reference's as long as it
 * is conservative.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FLAVOUR_WRAP 0  /* LinkCell:  freud/locality/LinkCell.cc:522 */
#define FLAVOUR_IMAGE 1 /* AABBQuery: freud/locality/AABBQuery.cc:93,125 */
#define FLAVOUR_GHOST 2 /* CellQuery: freud/locality/CellQuery.cc:107, CellIterator.h:167 */

typedef struct
{
    float Lx, Ly, Lz, xy, xz, yz;
    float lox, loy, loz;
    int is2d;
} box_t;

/* freud/box/Box.h:100-115 (setL) */
static box_t box_make(const float* b, int is2d)
{
    box_t x;
    x.Lx = b[0];
    x.Ly = b[1];
    x.Lz = is2d ? 0.0f : b[2];
    x.xy = b[3];
    x.xz = b[4];
    x.yz = b[5];
    x.is2d = is2d;
    /* m_hi = m_L / 2.0f  (vec / scalar multiplies by the reciprocal, VectorMath.h:208-212) */
    const float half = 1.0f / 2.0f;
    float hx = x.Lx * half, hy = x.Ly * half, hz = x.Lz * half;
    x.lox = -hx;
    x.loy = -hy;
    x.loz = -hz;
    return x;
}

/* freud/box/Box.h:243-255 */
static void box_fractional(const box_t* b, const float v[3], float f[3])
{
    float dx = v[0] - b->lox;
    float dy = v[1] - b->loy;
    float dz = v[2] - b->loz;
    float t0 = b->yz * b->xy;
    float t1 = b->xz - t0;
    float t2 = t1 * v[2];
    float t3 = b->xy * v[1];
    float t4 = t2 + t3;
    dx = dx - t4;
    float t5 = b->yz * v[2];
    dy = dy - t5;
    f[0] = dx / b->Lx;
    f[1] = dy / b->Ly;
    f[2] = dz / b->Lz; /* 0/0 -> NaN in 2-D, overwritten below */
    if (b->is2d)
    {
        f[2] = 0.0f;
    }
}

/* freud/box/Box.h:212-222 */
static void box_absolute(const box_t* b, const float f[3], float v[3])
{
    float px = f[0] * b->Lx;
    float py = f[1] * b->Ly;
    float pz = f[2] * b->Lz;
    float x = b->lox + px;
    float y = b->loy + py;
    float z = b->loz + pz;
    float a0 = b->xy * y;
    float a1 = b->xz * z;
    float a2 = a0 + a1;
    x = x + a2;
    float a3 = b->yz * z;
    y = y + a3;
    if (b->is2d)
    {
        z = 0.0f;
    }
    v[0] = x;
    v[1] = y;
    v[2] = z;
}

/* freud/util/utils.h:29-32 */
static float modulus_positive_one(float a)
{
    float t = fmodf(a, 1.0f);
    float u = t + 1.0f;
    return fmodf(u, 1.0f);
}

/* freud/box/Box.h:307-329 (all three axes periodic: queries in aperiodic boxes are rejected upstream,
 * NeighborQuery.h:133-138) */
static void box_wrap(const box_t* b, const float v[3], float out[3])
{
    float f[3];
    box_fractional(b, v, f);
    f[0] = modulus_positive_one(f[0]);
    f[1] = modulus_positive_one(f[1]);
    f[2] = modulus_positive_one(f[2]);
    box_absolute(b, f, out);
}

/* freud/box/Box.h:489-497 */
static void box_plane_distance(const box_t* b, float d[3])
{
    float t = b->xy * b->yz - b->xz;
    d[0] = b->Lx / sqrtf(1.0f + b->xy * b->xy + t * t);
    d[1] = b->Ly / sqrtf(1.0f + b->yz * b->yz);
    d[2] = b->Lz;
}

static float box_volume(const box_t* b)
{
    if (b->is2d)
    {
        return b->Lx * b->Ly;
    }
    return b->Lx * b->Ly * b->Lz;
}

/* dot: (x*x + y*y) + z*z, freud/util/VectorMath.h:270-273 */
static float dot3(const float r[3])
{
    float a = r[0] * r[0];
    float c = r[1] * r[1];
    float d = r[2] * r[2];
    float s = a + c;
    return s + d;
}

/* freud/locality/NeighborQuery.h:496-564: image k = i*a + j*b + k*c, image 0 first */
/* CellQuery::generateGhosts' displacement for a point seen across w boundaries (CellQuery.h:246-262): shift = 0,
 * then += +-a, += +-b, += +-c for the non-zero components (vec3 adds, component by component).  The query image
 * k of box_images corresponds to w = -k. */
static void ghost_shift(const box_t* b, const int w[3], float sh[3])
{
    float a[3] = {b->Lx, 0.0f, 0.0f};
    float bb[3] = {b->Ly * b->xy, b->Ly, 0.0f};
    float c[3] = {b->Lz * b->xz, b->Lz * b->yz, b->Lz};
    const float* vec[3] = {a, bb, c};
    sh[0] = sh[1] = sh[2] = 0.0f;
    for (int d = 0; d < 3; ++d)
    {
        if (w[d] != 0)
        {
            for (int e = 0; e < 3; ++e)
            {
                float v = w[d] > 0 ? vec[d][e] : -vec[d][e];
                sh[e] = sh[e] + v;
            }
        }
    }
}

static int box_images(const box_t* b, float img[27][3], int ijk[27][3])
{
    float a[3] = {b->Lx, 0.0f, 0.0f};
    float bb[3] = {b->Ly * b->xy, b->Ly, 0.0f};
    float c[3] = {0.0f, 0.0f, 0.0f};
    if (!b->is2d)
    {
        c[0] = b->Lz * b->xz;
        c[1] = b->Lz * b->yz;
        c[2] = b->Lz;
    }
    int n = 0;
    img[0][0] = img[0][1] = img[0][2] = 0.0f;
    ijk[0][0] = ijk[0][1] = ijk[0][2] = 0;
    n = 1;
    for (int i = -1; i <= 1; ++i)
    {
        for (int j = -1; j <= 1; ++j)
        {
            for (int k = -1; k <= 1; ++k)
            {
                if (i == 0 && j == 0 && k == 0)
                {
                    continue;
                }
                if (k != 0 && b->is2d)
                {
                    continue;
                }
                for (int d = 0; d < 3; ++d)
                {
                    /* float(i)*a + float(j)*b + float(k)*c, left to right */
                    float ta = (float) i * a[d];
                    float tb = (float) j * bb[d];
                    float tc = (float) k * c[d];
                    float s = ta + tb;
                    img[n][d] = s + tc;
                }
                ijk[n][0] = i;
                ijk[n][1] = j;
                ijk[n][2] = k;
                ++n;
            }
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * Conservative candidate grid (NOT the reference's cell list: see E1-E4 above).
 * ---------------------------------------------------------------------------------------------- */
typedef struct
{
    int dim[3];
    uint32_t* start; /* ncell + 1 */
    uint32_t* order; /* point indices grouped by cell */
    int* cell;       /* per point: cell coordinate packed */
    int* shift;      /* per point: 3 integer image offsets floor(frac) */
    size_t ncell;
} grid_t;

static void point_cell(const box_t* b, const int dim[3], const float p[3], int c[3], int n[3])
{
    float f[3];
    float v[3] = {p[0], p[1], b->is2d ? 0.0f : p[2]};
    box_fractional(b, v, f);
    for (int d = 0; d < 3; ++d)
    {
        float fl = floorf(f[d]);
        if (!(fabsf(fl) < 1.0e6f))
        {
            fl = 0.0f; /* NaN/inf guard; such points cannot be neighbours of anything */
        }
        float r = f[d] - fl;
        int ci = (int) (r * (float) dim[d]);
        if (ci >= dim[d])
        {
            ci = dim[d] - 1;
        }
        if (ci < 0)
        {
            ci = 0;
        }
        c[d] = ci;
        n[d] = (int) fl;
    }
}

static void grid_dims(const box_t* b, float r_search, size_t n_points, int dim[3])
{
    float pd[3];
    box_plane_distance(b, pd);
    double lmax = fmax(fmax(b->Lx, b->Ly), b->Lz);
    /* cell width > r_search by a margin that covers float32 rounding of coordinates of size ~L */
    double w = (double) r_search * (1.0 + 1.0e-4) + 1.0e-5 * lmax;
    for (int d = 0; d < 3; ++d)
    {
        double q = (double) pd[d] / w;
        int v = (int) floor(q);
        if (v < 1)
        {
            v = 1;
        }
        dim[d] = v;
    }
    if (b->is2d)
    {
        dim[2] = 1;
    }
    /* keep the grid O(N) cells */
    double cap = 4.0 * (double) n_points + 64.0;
    while ((double) dim[0] * dim[1] * dim[2] > cap)
    {
        for (int d = 0; d < 3; ++d)
        {
            if (dim[d] > 1)
            {
                dim[d] = (dim[d] + 1) / 2;
            }
        }
    }
}

static int grid_build(grid_t* g, const box_t* b, float r_search, const float* pts, size_t n)
{
    grid_dims(b, r_search, n, g->dim);
    g->ncell = (size_t) g->dim[0] * g->dim[1] * g->dim[2];
    g->start = (uint32_t*) calloc(g->ncell + 1, sizeof(uint32_t));
    g->order = (uint32_t*) malloc((n ? n : 1) * sizeof(uint32_t));
    g->cell = (int*) malloc((n ? n : 1) * sizeof(int));
    g->shift = (int*) malloc((n ? n : 1) * 3 * sizeof(int));
    if (!g->start || !g->order || !g->cell || !g->shift)
    {
        return -1;
    }
    for (size_t i = 0; i < n; ++i)
    {
        int c[3];
        point_cell(b, g->dim, pts + 3 * i, c, g->shift + 3 * i);
        int idx = (c[2] * g->dim[1] + c[1]) * g->dim[0] + c[0];
        g->cell[i] = idx;
        g->start[idx + 1]++;
    }
    for (size_t c = 0; c < g->ncell; ++c)
    {
        g->start[c + 1] += g->start[c];
    }
    uint32_t* cursor = (uint32_t*) malloc((g->ncell ? g->ncell : 1) * sizeof(uint32_t));
    if (!cursor)
    {
        return -1;
    }
    memcpy(cursor, g->start, g->ncell * sizeof(uint32_t));
    for (size_t i = 0; i < n; ++i)
    {
        g->order[cursor[g->cell[i]]++] = (uint32_t) i; /* ascending index inside each cell */
    }
    free(cursor);
    return 0;
}

static void grid_free(grid_t* g)
{
    free(g->start);
    free(g->order);
    free(g->cell);
    free(g->shift);
}

/* Per axis: the list of (cell coordinate, wrap count) slots to visit around home cell c. */
static int axis_slots(int dim, int c, int cells[3], int wraps[3])
{
    if (dim >= 3)
    {
        for (int o = -1; o <= 1; ++o)
        {
            int t = c + o;
            int w = 0;
            if (t < 0)
            {
                t += dim;
                w = -1;
            }
            else if (t >= dim)
            {
                t -= dim;
                w = 1;
            }
            cells[o + 1] = t;
            wraps[o + 1] = w;
        }
        return 3;
    }
    for (int t = 0; t < dim; ++t)
    {
        cells[t] = t;
        wraps[t] = 2; /* 2 == "ambiguous: try all three images on this axis" */
    }
    return dim;
}

typedef struct
{
    uint32_t j;
    float d;
    float v[3];
} hit_t;

typedef struct
{
    hit_t* data;
    size_t size, cap;
} hitvec_t;

static int hit_push(hitvec_t* h, uint32_t j, float d, const float v[3])
{
    if (h->size == h->cap)
    {
        size_t nc = h->cap ? 2 * h->cap : 64;
        hit_t* nd = (hit_t*) realloc(h->data, nc * sizeof(hit_t));
        if (!nd)
        {
            return -1;
        }
        h->data = nd;
        h->cap = nc;
    }
    h->data[h->size].j = j;
    h->data[h->size].d = d;
    memcpy(h->data[h->size].v, v, 3 * sizeof(float));
    h->size++;
    return 0;
}

/* NeighborBond::less_as_tuple restricted to one row with weight == 1 (NeighborBond.h:80-95) */
static int cmp_hit_j(const void* a, const void* b)
{
    const hit_t* x = (const hit_t*) a;
    const hit_t* y = (const hit_t*) b;
    if (x->j != y->j)
    {
        return x->j < y->j ? -1 : 1;
    }
    if (x->d != y->d)
    {
        return x->d < y->d ? -1 : 1;
    }
    return 0;
}

/* NeighborBond::less_as_distance restricted to one row (NeighborBond.h:97-112) */
static int cmp_hit_d(const void* a, const void* b)
{
    const hit_t* x = (const hit_t*) a;
    const hit_t* y = (const hit_t*) b;
    if (x->d != y->d)
    {
        return x->d < y->d ? -1 : 1;
    }
    if (x->j != y->j)
    {
        return x->j < y->j ? -1 : 1;
    }
    return 0;
}

/* All ball-query hits of one query point, in candidate order.
 * wrap flavour:  LinkCell.cc:514-528 ; image flavour: AABBQuery.cc:77-150 */
static int ball_hits(const box_t* b, const grid_t* g, int flavour, const float* pts, const float q_in[3],
                     uint32_t qi, float r_max, float r_min, int exclude_ii, float img[27][3], int ijk[27][3],
                     int n_img, hitvec_t* out)
{
    float r_max_sq = r_max * r_max;
    float r_min_sq = r_min * r_min;
    float q[3] = {q_in[0], q_in[1], q_in[2]};
    if (flavour == FLAVOUR_IMAGE && b->is2d)
    {
        q[2] = 0.0f; /* AABBQuery.cc:84-87 */
    }
    int c[3], nq[3];
    point_cell(b, g->dim, q, c, nq);
    int cx[3], cy[3], cz[3], wx[3], wy[3], wz[3];
    int nx = axis_slots(g->dim[0], c[0], cx, wx);
    int ny = axis_slots(g->dim[1], c[1], cy, wy);
    int nz = axis_slots(g->dim[2], c[2], cz, wz);
    for (int iz = 0; iz < nz; ++iz)
    {
        for (int iy = 0; iy < ny; ++iy)
        {
            for (int ix = 0; ix < nx; ++ix)
            {
                size_t cell = ((size_t) cz[iz] * g->dim[1] + cy[iy]) * g->dim[0] + cx[ix];
                for (uint32_t s = g->start[cell]; s < g->start[cell + 1]; ++s)
                {
                    uint32_t j = g->order[s];
                    if (exclude_ii && j == qi)
                    {
                        continue;
                    }
                    const float* pj = pts + 3 * (size_t) j;
                    if (flavour == FLAVOUR_WRAP)
                    {
                        float dlt[3] = {pj[0] - q[0], pj[1] - q[1], pj[2] - q[2]};
                        float r[3];
                        box_wrap(b, dlt, r);
                        float r_sq = dot3(r);
                        if (r_sq < r_max_sq && r_sq >= r_min_sq)
                        {
                            if (hit_push(out, j, sqrtf(r_sq), r))
                            {
                                return -1;
                            }
                        }
                    }
                    else
                    {
                        /* AABBQuery.cc:118-122 zeroes z in 2-D boxes; CellQuery takes the points as they are */
                        float p[3] = {pj[0], pj[1], (flavour == FLAVOUR_IMAGE && b->is2d) ? 0.0f : pj[2]};
                        const int* nj = g->shift + 3 * (size_t) j;
                        int w[3] = {wx[ix], wy[iy], wz[iz]};
                        for (int k = 0; k < n_img; ++k)
                        {
                            int ok = 1;
                            for (int d = 0; d < 3; ++d)
                            {
                                if (w[d] != 2 && ijk[k][d] != nj[d] - nq[d] - w[d])
                                {
                                    ok = 0;
                                }
                            }
                            if (!ok)
                            {
                                continue;
                            }
                            float r[3];
                            if (flavour == FLAVOUR_GHOST)
                            {
                                /* ghost = point + shift (CellQuery.cc:107), a real point is stored as it is (:121);
                                 * delta = ghost - query (CellIterator.h:167) */
                                int wk[3] = {-ijk[k][0], -ijk[k][1], -ijk[k][2]};
                                float sh[3];
                                ghost_shift(b, wk, sh);
                                for (int d = 0; d < 3; ++d)
                                {
                                    float gpos = k == 0 ? p[d] : p[d] + sh[d];
                                    r[d] = gpos - q[d];
                                }
                            }
                            else
                            {
                                float qk[3] = {q[0] + img[k][0], q[1] + img[k][1], q[2] + img[k][2]};
                                r[0] = p[0] - qk[0];
                                r[1] = p[1] - qk[1];
                                r[2] = p[2] - qk[2];
                            }
                            float r_sq = dot3(r);
                            if (r_sq < r_max_sq && r_sq >= r_min_sq)
                            {
                                if (hit_push(out, j, sqrtf(r_sq), r))
                                {
                                    return -1;
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Public entry points
 * ---------------------------------------------------------------------------------------------- */
typedef struct
{
    uint64_t n_bonds;
    uint32_t n_query;
    uint32_t* neighbors; /* n_bonds x 2 */
    float* distances;
    float* weights;
    float* vectors; /* n_bonds x 3 */
    uint32_t* segments;
    uint32_t* counts;
} port_nlist_t;

void fport_nlist_free(port_nlist_t* nl)
{
    if (!nl)
    {
        return;
    }
    free(nl->neighbors);
    free(nl->distances);
    free(nl->weights);
    free(nl->vectors);
    free(nl->segments);
    free(nl->counts);
    free(nl);
}

uint64_t fport_nlist_size(const port_nlist_t* nl)
{
    return nl->n_bonds;
}

void fport_nlist_copy(const port_nlist_t* nl, uint32_t* neighbors, float* distances, float* weights,
                      float* vectors, uint32_t* segments, uint32_t* counts)
{
    memcpy(neighbors, nl->neighbors, nl->n_bonds * 2 * sizeof(uint32_t));
    memcpy(distances, nl->distances, nl->n_bonds * sizeof(float));
    memcpy(weights, nl->weights, nl->n_bonds * sizeof(float));
    memcpy(vectors, nl->vectors, nl->n_bonds * 3 * sizeof(float));
    memcpy(segments, nl->segments, nl->n_query * sizeof(uint32_t));
    memcpy(counts, nl->counts, nl->n_query * sizeof(uint32_t));
}

/* rows -> NeighborList SoA; segments of empty rows stay 0 (NeighborList.cc:199-232) */
static port_nlist_t* assemble(hitvec_t* rows, uint32_t nq)
{
    port_nlist_t* nl = (port_nlist_t*) calloc(1, sizeof(port_nlist_t));
    uint64_t total = 0;
    for (uint32_t i = 0; i < nq; ++i)
    {
        total += rows[i].size;
    }
    nl->n_bonds = total;
    nl->n_query = nq;
    size_t t = total ? total : 1;
    nl->neighbors = (uint32_t*) malloc(t * 2 * sizeof(uint32_t));
    nl->distances = (float*) malloc(t * sizeof(float));
    nl->weights = (float*) malloc(t * sizeof(float));
    nl->vectors = (float*) malloc(t * 3 * sizeof(float));
    nl->segments = (uint32_t*) calloc(nq ? nq : 1, sizeof(uint32_t));
    nl->counts = (uint32_t*) calloc(nq ? nq : 1, sizeof(uint32_t));
    uint64_t off = 0;
    for (uint32_t i = 0; i < nq; ++i)
    {
        if (rows[i].size)
        {
            nl->segments[i] = (uint32_t) off;
            nl->counts[i] = (uint32_t) rows[i].size;
        }
        for (size_t k = 0; k < rows[i].size; ++k, ++off)
        {
            nl->neighbors[2 * off] = i;
            nl->neighbors[2 * off + 1] = rows[i].data[k].j;
            nl->distances[off] = rows[i].data[k].d;
            nl->weights[off] = 1.0f;
            memcpy(nl->vectors + 3 * off, rows[i].data[k].v, 3 * sizeof(float));
        }
        free(rows[i].data);
    }
    free(rows);
    return nl;
}

/* query(ball).toNeighborList(sort_by_distance): NeighborQuery.h:434-481 */
port_nlist_t* fport_ball_nlist(int flavour, const float* box6, int is2d, const float* pts, uint32_t n,
                               const float* qpts, uint32_t nq, float r_max, float r_min, int exclude_ii,
                               int sort_by_distance)
{
    box_t b = box_make(box6, is2d);
    grid_t g;
    if (grid_build(&g, &b, r_max, pts, n))
    {
        return NULL;
    }
    float img[27][3];
    int ijk[27][3];
    int n_img = box_images(&b, img, ijk);
    hitvec_t* rows = (hitvec_t*) calloc(nq ? nq : 1, sizeof(hitvec_t));
    int failed = 0;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t) nq; ++i)
    {
        if (ball_hits(&b, &g, flavour, pts, qpts + 3 * i, (uint32_t) i, r_max, r_min, exclude_ii, img, ijk, n_img,
                      &rows[i]))
        {
            failed = 1;
        }
        qsort(rows[i].data, rows[i].size, sizeof(hit_t), sort_by_distance ? cmp_hit_d : cmp_hit_j);
    }
    grid_free(&g);
    if (failed)
    {
        return NULL;
    }
    return assemble(rows, nq);
}

/* RegularAxis::bin, freud/util/Histogram.h:152-174 (returns -1 for the overflow bin) */
static int64_t axis_bin(float value, float mn, float mx, float inv_width, uint32_t nbins)
{
    if (value < mn || value >= mx)
    {
        return -1;
    }
    float d = value - mn;
    float val = d * inv_width;
    int64_t bin = (int64_t) val; /* truncation, == _mm_cvtt_ss2si for in-range values */
    if ((uint64_t) bin == nbins)
    {
        return bin - 1;
    }
    return bin;
}

static void axis_params(uint32_t bins, float mn, float mx, float* width, float* inv)
{
    float span = mx - mn;
    *width = span / (float) bins; /* Histogram.h:129 */
    *inv = 1.0f / *width;         /* Histogram.h:130 */
}

/* RDF::accumulate without a NeighborList (RDF.cc:101-110 via NeighborComputeFunctional.h:195-217):
 * adds this frame's counts into counts[bins] (u32, wrapping like the reference's unsigned int). */
int fport_rdf_accumulate(int flavour, const float* box6, int is2d, const float* pts, uint32_t n, const float* qpts,
                         uint32_t nq, float q_r_max, float q_r_min, int exclude_ii, uint32_t bins, float bin_r_min,
                         float bin_r_max, uint32_t* counts)
{
    box_t b = box_make(box6, is2d);
    grid_t g;
    if (grid_build(&g, &b, q_r_max, pts, n))
    {
        return -1;
    }
    float img[27][3];
    int ijk[27][3];
    int n_img = box_images(&b, img, ijk);
    float width, inv;
    axis_params(bins, bin_r_min, bin_r_max, &width, &inv);
    int failed = 0;
#pragma omp parallel
    {
        uint32_t* local = (uint32_t*) calloc(bins, sizeof(uint32_t));
        hitvec_t h = {0, 0, 0};
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t) nq; ++i)
        {
            h.size = 0;
            if (ball_hits(&b, &g, flavour, pts, qpts + 3 * i, (uint32_t) i, q_r_max, q_r_min, exclude_ii, img, ijk,
                          n_img, &h))
            {
                failed = 1;
            }
            for (size_t k = 0; k < h.size; ++k)
            {
                int64_t bin = axis_bin(h.data[k].d, bin_r_min, bin_r_max, inv, bins);
                if (bin >= 0)
                {
                    local[bin]++;
                }
            }
        }
#pragma omp critical
        for (uint32_t k = 0; k < bins; ++k)
        {
            counts[k] += local[k];
        }
        free(local);
        free(h.data);
    }
    grid_free(&g);
    return failed ? -1 : 0;
}

/* RDF from a NeighborList: one increment per stored distance (NeighborComputeFunctional.h:180-193) */
void fport_rdf_accumulate_distances(const float* distances, uint64_t n_bonds, uint32_t bins, float bin_r_min,
                                    float bin_r_max, uint32_t* counts)
{
    float width, inv;
    axis_params(bins, bin_r_min, bin_r_max, &width, &inv);
    for (uint64_t k = 0; k < n_bonds; ++k)
    {
        int64_t bin = axis_bin(distances[k], bin_r_min, bin_r_max, inv, bins);
        if (bin >= 0)
        {
            counts[bin]++;
        }
    }
}

/* RDF::RDF volumes (RDF.cc:25-64) + RDF::reduce (RDF.cc:73-99) + bin edges/centres (Histogram.h:126-138,
 * Axis::getBinCenters). */
void fport_rdf_reduce(const uint32_t* counts, uint32_t bins, float r_min, float r_max, const float* box6, int is2d,
                      uint32_t n_points, uint32_t n_query_points, uint32_t frames, int finite_size, float* g_r,
                      float* n_r, float* edges, float* centers)
{
    box_t b = box_make(box6, is2d);
    float width, inv;
    axis_params(bins, r_min, r_max, &width, &inv);
    float* e = (float*) malloc((bins + 1) * sizeof(float));
    for (uint32_t i = 0; i <= bins; ++i)
    {
        float t = (float) i * width;
        e[i] = r_min + t;
    }
    float nqp = (float) n_query_points;
    float number_density = nqp / box_volume(&b);
    if (finite_size)
    {
        float ratio = (float) (n_query_points - 1) / (float) n_query_points;
        number_density = number_density * ratio;
    }
    float np = (float) n_points;
    float nf = (float) frames;
    float den = np * number_density;
    den = den * nf;
    float prefactor = 1.0f / den;
    /* volume_prefactor = (4.0f/3.0f) * M_PI : float * double, stored to float */
    const float volume_prefactor = (float) ((double) (4.0f / 3.0f) * M_PI);
    for (uint32_t i = 0; i < bins; ++i)
    {
        float r = e[i];
        float nextr = e[i + 1];
        float vol;
        if (is2d)
        {
            /* M_PI * (nextr*nextr - r*r): float difference promoted to double, product stored to float */
            float a = nextr * nextr;
            float c = r * r;
            float diff = a - c;
            vol = (float) (M_PI * (double) diff);
        }
        else
        {
            float a = nextr * nextr;
            a = a * nextr;
            float c = r * r;
            c = c * r;
            float diff = a - c;
            vol = volume_prefactor * diff;
        }
        float t = (float) counts[i] * prefactor;
        g_r[i] = t / vol;
    }
    float pre2 = 1.0f / (nqp * (float) frames);
    n_r[0] = (float) counts[0] * pre2;
    for (uint32_t i = 1; i < bins; ++i)
    {
        float t = (float) counts[i] * pre2;
        n_r[i] = n_r[i - 1] + t;
    }
    if (edges)
    {
        memcpy(edges, e, (bins + 1) * sizeof(float));
    }
    if (centers)
    {
        for (uint32_t i = 0; i < bins; ++i)
        {
            float s = e[i] + e[i + 1];
            centers[i] = s / 2.0f; /* Histogram.h Axis::getBinCenters: (edge[i] + edge[i+1]) / 2 */
        }
    }
    free(e);
}

/* ---- kNN (E3), image flavour ------------------------------------------------------------------- */
/* For each query: the num_neighbors smallest closest-image distances (AABBQuery.cc:152-281), d >= r_min,
 * d <= r_max.  Brute force over the points; used for small and medium N only. */
port_nlist_t* fport_knn_nlist(const float* box6, int is2d, const float* pts, uint32_t n, const float* qpts,
                              uint32_t nq, uint32_t k, float r_max, float r_min, int exclude_ii,
                              int sort_by_distance)
{
    box_t b = box_make(box6, is2d);
    float img[27][3];
    int ijk[27][3];
    int n_img = box_images(&b, img, ijk);
    hitvec_t* rows = (hitvec_t*) calloc(nq ? nq : 1, sizeof(hitvec_t));
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < (int64_t) nq; ++i)
    {
        hitvec_t all = {0, 0, 0};
        float q[3] = {qpts[3 * i], qpts[3 * i + 1], is2d ? 0.0f : qpts[3 * i + 2]};
        for (uint32_t j = 0; j < n; ++j)
        {
            if (exclude_ii && j == (uint32_t) i)
            {
                continue;
            }
            float p[3] = {pts[3 * (size_t) j], pts[3 * (size_t) j + 1], is2d ? 0.0f : pts[3 * (size_t) j + 2]};
            float best = INFINITY, bv[3] = {0, 0, 0};
            for (int m = 0; m < n_img; ++m)
            {
                float qk[3] = {q[0] + img[m][0], q[1] + img[m][1], q[2] + img[m][2]};
                float r[3] = {p[0] - qk[0], p[1] - qk[1], p[2] - qk[2]};
                float d = sqrtf(dot3(r));
                if (d < best)
                {
                    best = d;
                    memcpy(bv, r, sizeof(bv));
                }
            }
            if (best >= r_min)
            {
                hit_push(&all, j, best, bv);
            }
        }
        qsort(all.data, all.size, sizeof(hit_t), cmp_hit_d);
        size_t keep = 0;
        while (keep < all.size && keep < k && !(all.data[keep].d > r_max))
        {
            ++keep;
        }
        all.size = keep;
        if (!sort_by_distance)
        {
            qsort(all.data, all.size, sizeof(hit_t), cmp_hit_j);
        }
        rows[i] = all;
    }
    return assemble(rows, nq);
}

/* ---- kNN, wrap flavour ------------------------------------------------------------------------- */
/* LinkCellQueryIterator::next (LinkCell.cc:575-679): every point once with r = Box::wrap(p_j - q), kept if
 * r_min^2 <= r.r < r_max^2 (:577-578, :617-622), sorted by distance, the first num_neighbors returned (:663-672).
 * The shell-by-shell early exit (:654-661) only stops once the k-th distance is inside the searched shells, so the
 * cell structure does not influence the answer.  Brute force over the points. */
port_nlist_t* fport_knn_nlist_wrap(const float* box6, int is2d, const float* pts, uint32_t n, const float* qpts,
                                   uint32_t nq, uint32_t k, float r_max, float r_min, int exclude_ii,
                                   int sort_by_distance)
{
    box_t b = box_make(box6, is2d);
    volatile float r_max_sq_v = r_max * r_max, r_min_sq_v = r_min * r_min;
    float r_max_sq = r_max_sq_v, r_min_sq = r_min_sq_v;
    hitvec_t* rows = (hitvec_t*) calloc(nq ? nq : 1, sizeof(hitvec_t));
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < (int64_t) nq; ++i)
    {
        hitvec_t all = {0, 0, 0};
        for (uint32_t j = 0; j < n; ++j)
        {
            if (exclude_ii && j == (uint32_t) i)
            {
                continue;
            }
            float dlt[3] = {pts[3 * (size_t) j] - qpts[3 * i], pts[3 * (size_t) j + 1] - qpts[3 * i + 1],
                            pts[3 * (size_t) j + 2] - qpts[3 * i + 2]};
            float r[3];
            box_wrap(&b, dlt, r);
            float r_sq = dot3(r);
            if (r_sq < r_max_sq && r_sq >= r_min_sq)
            {
                hit_push(&all, j, sqrtf(r_sq), r);
            }
        }
        qsort(all.data, all.size, sizeof(hit_t), cmp_hit_d);
        if (all.size > k)
        {
            all.size = k;
        }
        if (!sort_by_distance)
        {
            qsort(all.data, all.size, sizeof(hit_t), cmp_hit_j);
        }
        rows[i] = all;
    }
    return assemble(rows, nq);
}

/* ---- Steinhardt ------------------------------------------------------------------------------- */
/* fsph::PointSPHEvaluator<float> (extern/fsph/src/spherical_harmonics.hpp:155-290), lmax <= 32 */
#define SPH_LMAX 32
typedef struct
{
    unsigned lmax;
    float pref[2 * (SPH_LMAX + 1) * SPH_LMAX];
    float sinpow[SPH_LMAX + 1];
    float th_re[SPH_LMAX + 1], th_im[SPH_LMAX + 1];
    float jac[(SPH_LMAX + 1) * (SPH_LMAX + 1)];
} sph_t;

static void sph_init(sph_t* s, unsigned lmax)
{
    s->lmax = lmax;
    unsigned f1 = lmax * (lmax + 1);
    for (unsigned m = 0; m < lmax + 1; ++m)
    {
        for (unsigned l = 1; l < lmax + 1; ++l)
        {
            s->pref[lmax * m + (l - 1)] = (float) (2 * sqrt(1 + (m - 0.5) / l) * sqrt(1 - (m - 0.5) / (l + 2 * m)));
        }
    }
    for (unsigned m = 0; m < lmax + 1; ++m)
    {
        if (lmax > 0)
        {
            s->pref[f1 + lmax * m + 0] = 0;
        }
        for (unsigned l = 2; l < lmax + 1; ++l)
        {
            s->pref[f1 + lmax * m + (l - 1)]
                = (float) (-sqrt(1.0 + 4.0 / (2 * l + 2 * m - 3)) * sqrt(1 - 1.0 / l) * sqrt(1.0 - 1.0 / (l + 2 * m)));
        }
    }
}

/* compute(phi = polar, theta = azimuth) */
static void sph_compute(sph_t* s, float phi, float theta)
{
    unsigned lmax = s->lmax;
    unsigned f1 = lmax * (lmax + 1);
    float sphi = sinf(phi);
    s->sinpow[0] = 1;
    for (unsigned i = 1; i < lmax + 1; ++i)
    {
        s->sinpow[i] = s->sinpow[i - 1] * sphi;
    }
    for (unsigned i = 0; i < lmax + 1; ++i)
    {
        float a = (float) i * theta;
        s->th_re[i] = cosf(a); /* exp(complex<float>(0, a)) */
        s->th_im[i] = sinf(a);
    }
    float cphi = cosf(phi);
    unsigned w = lmax + 1;
    for (unsigned m = 0; m < lmax + 1; ++m)
    {
        if (m > 0)
        {
            s->jac[w * m] = (float) (s->jac[w * (m - 1)] * sqrt(1 + 1.0 / 2 / m));
        }
        else
        {
            s->jac[0] = (float) (1 / sqrt(2));
        }
        if (lmax > 0)
        {
            float t = cphi * s->pref[lmax * m + 0];
            s->jac[w * m + 1] = t * s->jac[w * m];
        }
        for (unsigned l = 2; l < lmax + 1; ++l)
        {
            float t = cphi * s->pref[lmax * m + (l - 1)];
            float a = t * s->jac[w * m + l - 1];
            float c = s->pref[f1 + lmax * m + (l - 1)] * s->jac[w * m + l - 2];
            s->jac[w * m + l] = a + c;
        }
    }
}

/* Y_lm for one l in Steinhardt's order m = 0..l, -1..-l with the Condon-Shortley sign on odd positive m
 * (Steinhardt.cc:31-52; iterator::operator* spherical_harmonics.hpp:78-93) */
static void sph_ylm(const sph_t* s, unsigned l, float* re, float* im)
{
    unsigned w = s->lmax + 1;
    for (unsigned k = 0; k < 2 * l + 1; ++k)
    {
        unsigned m = k <= l ? k : k - l;
        float legendre = s->sinpow[m] * s->jac[w * m + (l - m)];
        float a = (float) (legendre / sqrt(2 * M_PI));
        float yr = a * s->th_re[m];
        float yi = a * s->th_im[m];
        if (k > l)
        {
            yi = -yi; /* conj */
        }
        float phase = (k <= l && k % 2 == 1) ? -1.0f : 1.0f;
        re[k] = phase * yr;
        im[k] = phase * yi;
    }
}

static float clampf(float v, float lo, float hi)
{
    return fmaxf(lo, fminf(v, hi));
}

/* Steinhardt::baseCompute for plain q_l (Steinhardt.cc:120-222) over a NeighborList given as CSR rows
 * (segments/counts as produced above), using the list's distances and recomputing delta with wrap.
 * ql: N x n_ls ; qlm: per l concatenated, N x (2l+1) complex (re,im interleaved), laid out l-major;
 * sys_qlm: sum_i qlm_i / N per l (float accumulation in index order). */
int fport_steinhardt(const float* box6, int is2d, const float* pts, uint32_t n, const uint32_t* nl_j,
                     const float* nl_d, const float* nl_w, const uint32_t* segments, const uint32_t* counts,
                     const uint32_t* ls, uint32_t n_ls, int weighted, float* ql, float* qlm, float* sys_qlm,
                     float* order)
{
    box_t b = box_make(box6, is2d);
    unsigned lmax = 0;
    size_t tot_m = 0;
    for (uint32_t a = 0; a < n_ls; ++a)
    {
        if (ls[a] > lmax)
        {
            lmax = ls[a];
        }
        tot_m += 2 * ls[a] + 1;
    }
    if (lmax > SPH_LMAX)
    {
        return -1;
    }
    size_t* l_off = (size_t*) malloc(n_ls * sizeof(size_t));
    size_t acc = 0;
    for (uint32_t a = 0; a < n_ls; ++a)
    {
        l_off[a] = acc;
        acc += (size_t) n * (2 * ls[a] + 1) * 2;
    }
    memset(qlm, 0, acc * sizeof(float));
    memset(ql, 0, (size_t) n * n_ls * sizeof(float));
#pragma omp parallel
    {
        sph_t s;
        sph_init(&s, lmax);
        float yr[2 * SPH_LMAX + 1], yi[2 * SPH_LMAX + 1];
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t) n; ++i)
        {
            float total_weight = 0;
            const float* ref = pts + 3 * i;
            uint32_t beg = counts[i] ? segments[i] : 0;
            for (uint32_t kb = 0; kb < counts[i]; ++kb)
            {
                uint32_t bond = beg + kb;
                const float* pj = pts + 3 * (size_t) nl_j[bond];
                float dlt[3] = {pj[0] - ref[0], pj[1] - ref[1], pj[2] - ref[2]};
                float delta[3];
                box_wrap(&b, dlt, delta);
                float weight = weighted ? nl_w[bond] : 1.0f;
                float phi = atan2f(delta[1], delta[0]);
                float dist = nl_d[bond];
                float theta = acosf(clampf(delta[2] / dist, -1.0f, 1.0f));
                if (dist == 0.0f)
                {
                    theta = 0;
                }
                sph_compute(&s, theta, phi);
                for (uint32_t a = 0; a < n_ls; ++a)
                {
                    unsigned nm = 2 * ls[a] + 1;
                    sph_ylm(&s, ls[a], yr, yi);
                    float* dst = qlm + l_off[a] + (size_t) i * nm * 2;
                    for (unsigned k = 0; k < nm; ++k)
                    {
                        float tr = weight * yr[k];
                        float ti = weight * yi[k];
                        dst[2 * k] = dst[2 * k] + tr;
                        dst[2 * k + 1] = dst[2 * k + 1] + ti;
                    }
                }
                total_weight = total_weight + weight;
            }
            for (uint32_t a = 0; a < n_ls; ++a)
            {
                unsigned nm = 2 * ls[a] + 1;
                float normalizationfactor = (float) (4.0 * M_PI / nm);
                float* dst = qlm + l_off[a] + (size_t) i * nm * 2;
                float sum = 0;
                for (unsigned k = 0; k < nm; ++k)
                {
                    dst[2 * k] = dst[2 * k] / total_weight;
                    dst[2 * k + 1] = dst[2 * k + 1] / total_weight;
                    float rr = dst[2 * k] * dst[2 * k];
                    float ii = dst[2 * k + 1] * dst[2 * k + 1];
                    float nn = rr + ii; /* std::norm */
                    sum = sum + nn;
                }
                sum = sum * normalizationfactor;
                ql[(size_t) i * n_ls + a] = sqrtf(sum);
            }
        }
    }
    /* system q_lm: the reference sums thread-local partials in scheduler order (non-deterministic);
     * here plain index order.  normalizeSystem: Steinhardt.cc:291-327 */
    size_t so = 0;
    for (uint32_t a = 0; a < n_ls; ++a)
    {
        unsigned nm = 2 * ls[a] + 1;
        float calc_norm = 0;
        for (unsigned k = 0; k < nm; ++k)
        {
            float sr = 0, si = 0;
            for (uint32_t i = 0; i < n; ++i)
            {
                const float* src = qlm + l_off[a] + (size_t) i * nm * 2;
                sr = sr + src[2 * k] / (float) n;
                si = si + src[2 * k + 1] / (float) n;
            }
            sys_qlm[so + 2 * k] = sr;
            sys_qlm[so + 2 * k + 1] = si;
            float nn = sr * sr + si * si;
            calc_norm = calc_norm + nn;
        }
        float normalizationfactor = (float) (4.0 * M_PI / nm);
        order[a] = sqrtf(calc_norm * normalizationfactor);
        so += 2 * nm;
    }
    (void) tot_m;
    free(l_off);
    return 0;
}

/* ---- small helpers exposed for the Box known-answer tests ---------------------------------------- */
/* op: 0 wrap, 1 makeFractional, 2 makeAbsolute */
void fport_box_apply(const float* box6, int is2d, int op, const float* in, uint32_t n, float* out)
{
    box_t b = box_make(box6, is2d);
    for (uint32_t i = 0; i < n; ++i)
    {
        if (op == 0)
        {
            box_wrap(&b, in + 3 * i, out + 3 * i);
        }
        else if (op == 1)
        {
            box_fractional(&b, in + 3 * i, out + 3 * i);
        }
        else
        {
            box_absolute(&b, in + 3 * i, out + 3 * i);
        }
    }
}

void fport_box_info(const float* box6, int is2d, float* volume, float* plane_dist3)
{
    box_t b = box_make(box6, is2d);
    *volume = box_volume(&b);
    box_plane_distance(&b, plane_dist3);
}

/* Number of candidate pair evaluations a 27-cell scheme with this grid performs ("pair evals" in bench.py):
 * sum over query points of the occupancy of the visited cells (self included). */
uint64_t fport_count_candidates(const float* box6, int is2d, const float* pts, uint32_t n, const float* qpts,
                                uint32_t nq, float r_max)
{
    box_t b = box_make(box6, is2d);
    grid_t g;
    if (grid_build(&g, &b, r_max, pts, n))
    {
        return 0;
    }
    uint64_t total = 0;
#pragma omp parallel for reduction(+ : total) schedule(static)
    for (int64_t i = 0; i < (int64_t) nq; ++i)
    {
        int c[3], nn[3];
        point_cell(&b, g.dim, qpts + 3 * i, c, nn);
        int cx[3], cy[3], cz[3], wx[3], wy[3], wz[3];
        int nx = axis_slots(g.dim[0], c[0], cx, wx);
        int ny = axis_slots(g.dim[1], c[1], cy, wy);
        int nz = axis_slots(g.dim[2], c[2], cz, wz);
        for (int iz = 0; iz < nz; ++iz)
        {
            for (int iy = 0; iy < ny; ++iy)
            {
                for (int ix = 0; ix < nx; ++ix)
                {
                    size_t cell = ((size_t) cz[iz] * g.dim[1] + cy[iy]) * g.dim[0] + cx[ix];
                    total += g.start[cell + 1] - g.start[cell];
                }
            }
        }
    }
    grid_free(&g);
    return total;
}

/* ---- Steinhardt options: second-shell average and w_l ------------------------------------------------------
 * Steinhardt::computeAve (Steinhardt.cc:224-289), aggregatewl (:329-359) with reduceWigner3j (Wigner3j.cc:22-57)
 * and normalizeSystem (:291-327), applied to the q_lm(i) and q_l(i) that fport_steinhardt produced.
 * Upstream tabulates the Wigner 3j symbols (l l l; m1 m2 m3) for l <= 20 (Wigner3j.cc:59-5832); the table is a
 * third-party-style constant, so it is restated here by the published algorithm behind it -- Racah's formula,
 *   (l l l; m1 m2 m3) = (-1)^m3 sqrt( prod (l +- m_i)! / ((3l+1)! (l!)^3) ) sum_k (-1)^k C(l,k) C(l,l-m1-k) C(l,l+m2-k)
 * with the sum exact in 128-bit integers -- and checked against upstream's numbers (identical as float for every
 * l <= 20, tests/test_oracle_port.py). */
static void wigner3j_table(int l, float* out)
{
    long double fact[3 * 32 + 2];
    fact[0] = 1.0L;
    for (int k = 1; k < 3 * l + 2; ++k)
    {
        fact[k] = fact[k - 1] * (long double) k;
    }
    __int128 binom[33][33];
    for (int a = 0; a <= l; ++a)
    {
        for (int k = 0; k <= l; ++k)
        {
            binom[a][k] = 0;
        }
        binom[a][0] = 1;
        for (int k = 1; k <= a; ++k)
        {
            binom[a][k] = binom[a - 1][k - 1] + (k <= a - 1 ? binom[a - 1][k] : 0);
        }
    }
    size_t counter = 0;
    for (int m1 = -l; m1 <= l; ++m1)
    {
        int lo = -l - m1 > -l ? -l - m1 : -l, hi = l - m1 < l ? l - m1 : l;
        for (int m2 = lo; m2 <= hi; ++m2)
        {
            int m3 = -m1 - m2;
            __int128 sum = 0;
            for (int k = 0; k <= l; ++k)
            {
                int k2 = l - m1 - k, k3 = l + m2 - k;
                if (k2 < 0 || k2 > l || k3 < 0 || k3 > l)
                {
                    continue;
                }
                __int128 term = binom[l][k] * binom[l][k2] * binom[l][k3];
                sum += (k & 1) ? -term : term;
            }
            long double pref = sqrtl(fact[l + m1] * fact[l - m1] * fact[l + m2] * fact[l - m2] * fact[l + m3]
                                     * fact[l - m3] / (fact[3 * l + 1] * fact[l] * fact[l] * fact[l]));
            long double v = ((m3 & 1) ? -1.0L : 1.0L) * pref * (long double) sum;
            out[counter++] = (float) (double) v; /* upstream stores doubles and uses float(w), Wigner3j.cc:52 */
        }
    }
}

static size_t wigner3j_count(int l)
{
    return (size_t) (3 * l * l + 3 * l + 1);
}

/* reduceWigner3j, Wigner3j.cc:22-57: sum over the table of Re(float(w) * s[m1] * s[m2] * s[m3]), std::complex<float>
 * products left to right; source is interleaved (re, im) in the order m = 0..l, -1..-l */
static float reduce_wigner3j(const float* src, int l, const float* w3j)
{
    float result = 0;
    size_t counter = 0;
    for (int m1 = -l; m1 <= l; ++m1)
    {
        int i1 = m1 < 0 ? l - m1 : m1;
        int lo = -l - m1 > -l ? -l - m1 : -l, hi = l - m1 < l ? l - m1 : l;
        for (int m2 = lo; m2 <= hi; ++m2)
        {
            int i2 = m2 < 0 ? l - m2 : m2;
            int m3 = -m1 - m2;
            int i3 = m3 < 0 ? l - m3 : m3;
            float w = w3j[counter++];
            float ar = w * src[2 * i1], ai = w * src[2 * i1 + 1];
            float t1 = ar * src[2 * i2], t2 = ai * src[2 * i2 + 1];
            float t3 = ar * src[2 * i2 + 1], t4 = ai * src[2 * i2];
            float br = t1 - t2, bi = t3 + t4;
            float u1 = br * src[2 * i3], u2 = bi * src[2 * i3 + 1];
            float cr = u1 - u2;
            result = result + cr;
        }
    }
    return result;
}

/* qlm: per l blocks as fport_steinhardt lays them out (n x (2l+1) complex each); ql: n x n_ls.
 * ql_out: what getQl() returns (the averaged q_l with average); particle_order: getParticleOrder(); order: getOrder(). */
int fport_steinhardt_options(uint32_t n, const uint32_t* nl_j, const uint32_t* segments, const uint32_t* counts,
                             const uint32_t* ls, uint32_t n_ls, int average, int wl, int wl_normalize,
                             const float* qlm, const float* ql, float* ql_out, float* particle_order, float* order)
{
    size_t total = 0;
    for (uint32_t a = 0; a < n_ls; ++a)
    {
        if (ls[a] > SPH_LMAX || (wl && ls[a] > 20))
        {
            return -1; /* getWigner3j throws std::out_of_range beyond l = 20 */
        }
        total += (size_t) n * (2 * ls[a] + 1) * 2;
    }
    float* ave = average ? (float*) calloc(total ? total : 1, sizeof(float)) : NULL;
    float* ql_ave = average ? (float*) calloc((size_t) n * n_ls + 1, sizeof(float)) : NULL;
    size_t off = 0;
    for (uint32_t a = 0; a < n_ls; ++a)
    {
        int l = (int) ls[a];
        unsigned nm = 2 * ls[a] + 1;
        float nf = (float) (4.0 * M_PI / nm);
        const float* blk = qlm + off;
        if (average)
        {
            float* ablk = ave + off;
            for (uint32_t i = 0; i < n; ++i)
            {
                float* dst = ablk + (size_t) i * nm * 2;
                uint32_t beg = counts[i] ? segments[i] : 0;
                unsigned neighborcount = 1; /* Steinhardt.cc:244 */
                for (uint32_t kb = 0; kb < counts[i]; ++kb)
                {
                    const float* src = blk + (size_t) nl_j[beg + kb] * nm * 2;
                    for (unsigned k = 0; k < 2 * nm; ++k)
                    {
                        dst[k] = dst[k] + src[k];
                    }
                    neighborcount++;
                }
                const float* own = blk + (size_t) i * nm * 2;
                float sum = 0;
                for (unsigned k = 0; k < nm; ++k)
                {
                    dst[2 * k] = dst[2 * k] + own[2 * k];
                    dst[2 * k + 1] = dst[2 * k + 1] + own[2 * k + 1];
                    dst[2 * k] = dst[2 * k] / (float) neighborcount;
                    dst[2 * k + 1] = dst[2 * k + 1] / (float) neighborcount;
                    float rr = dst[2 * k] * dst[2 * k], ii = dst[2 * k + 1] * dst[2 * k + 1];
                    float nn = rr + ii;
                    sum = sum + nn;
                }
                sum = sum * nf;
                ql_ave[(size_t) i * n_ls + a] = sqrtf(sum);
            }
        }
        const float* source = average ? ave + off : blk;
        const float* norm_src = average ? ql_ave : ql;
        float* w3j = NULL;
        if (wl)
        {
            w3j = (float*) malloc(wigner3j_count(l) * sizeof(float));
            wigner3j_table(l, w3j);
        }
        for (uint32_t i = 0; i < n; ++i)
        {
            float q = norm_src[(size_t) i * n_ls + a];
            ql_out[(size_t) i * n_ls + a] = q;
            float po = q;
            if (wl)
            {
                po = reduce_wigner3j(source + (size_t) i * nm * 2, l, w3j);
                if (wl_normalize)
                {
                    float normalization = sqrtf(nf) / q;
                    float n2 = normalization * normalization;
                    float n3 = n2 * normalization;
                    po = po * n3;
                }
            }
            particle_order[(size_t) i * n_ls + a] = po;
        }
        /* system q_lm (index order here; scheduler order upstream) and normalizeSystem */
        float sys[2 * (2 * SPH_LMAX + 1)];
        float calc_norm = 0;
        for (unsigned k = 0; k < nm; ++k)
        {
            float sr = 0, si = 0;
            for (uint32_t i = 0; i < n; ++i)
            {
                const float* src = source + (size_t) i * nm * 2;
                sr = sr + src[2 * k] / (float) n;
                si = si + src[2 * k + 1] / (float) n;
            }
            sys[2 * k] = sr;
            sys[2 * k + 1] = si;
            float nn = sr * sr + si * si;
            calc_norm = calc_norm + nn;
        }
        float ql_system = sqrtf(calc_norm * nf);
        if (wl)
        {
            float wl_system = reduce_wigner3j(sys, l, w3j);
            if (wl_normalize)
            {
                float normalization = sqrtf(nf) / ql_system;
                wl_system = wl_system * (normalization * normalization * normalization);
            }
            order[a] = wl_system;
        }
        else
        {
            order[a] = ql_system;
        }
        free(w3j);
        off += (size_t) n * nm * 2;
    }
    free(ave);
    free(ql_ave);
    return 0;
}

/* the table alone, for the test that compares it with upstream's numbers */
int fport_wigner3j(uint32_t l, float* out)
{
    if (l > 32)
    {
        return -1;
    }
    wigner3j_table((int) l, out);
    return (int) wigner3j_count((int) l);
}

/* ---- LocalDensity ------------------------------------------------------------------------------------
 * LocalDensity::compute (freud/density/LocalDensity.cc:38-84) over the rows of a NeighborList, in list order:
 * a point wholly inside r_max counts 1, one straddling it 1 + (r_max - (d + diameter/2)) / diameter (:58-69);
 * density = count / (M_PI r^2) in 2-D boxes (:48, double product rounded once) or / (float(4/3 pi) r r r) (:49).
 * Rows without bonds keep 0 (the arrays are zero-initialised and only written inside the bond loop). */
int fport_local_density(const float* nl_d, const uint32_t* segments, const uint32_t* counts, uint32_t n_query,
                        float r_max, float diameter, int is2d, float* num_neighbors, float* density)
{
    float area = (float) (M_PI * r_max * r_max);
    float volume = (float) (4.0 / 3.0 * M_PI);
    volume = volume * r_max;
    volume = volume * r_max;
    volume = volume * r_max;
    float half = diameter / 2.0f;
    float inner = r_max - half;
    for (uint32_t i = 0; i < n_query; ++i)
    {
        float num = 0;
        uint32_t beg = counts[i] ? segments[i] : 0;
        for (uint32_t kb = 0; kb < counts[i]; ++kb)
        {
            float d = nl_d[beg + kb];
            if (d < inner)
            {
                num = num + 1.0f;
            }
            else
            {
                float t = d + half;
                float u = r_max - t;
                float part = u / diameter;
                float inc = 1.0f + part;
                num = num + inc;
            }
        }
        num_neighbors[i] = num;
        density[i] = counts[i] ? (is2d ? num / area : num / volume) : 0.0f;
    }
    return 0;
}

/* ---- CorrelationFunction ---------------------------------------------------------------------------------
 * CorrelationFunction::accumulate + reduce (freud/density/CorrelationFunction.cc:49-59, 69-95) over the bonds of a
 * NeighborList: bin = RegularAxis(bins, 0, r_max).bin(distance) (overflow bonds are dropped, Histogram.h:313-320),
 * count[bin]++, sum[bin] += conj(values[j]) * query_values[i] in complex<double>; then sum / count where count != 0.
 * Bonds are added in list order here; upstream adds per-thread partial sums, so agreement is to double rounding. */
int fport_correlation(const uint32_t* nl_ij, const float* nl_d, uint64_t n_bonds, const double* values,
                      const double* query_values, uint32_t bins, float r_max, double* correlation, uint32_t* counts)
{
    volatile float width_v = r_max / (float) bins;
    volatile float inv_v = 1.0f / width_v;
    float inv = inv_v;
    memset(correlation, 0, 2 * (size_t) bins * sizeof(double));
    memset(counts, 0, (size_t) bins * sizeof(uint32_t));
    for (uint64_t k = 0; k < n_bonds; ++k)
    {
        int64_t bin = axis_bin(nl_d[k], 0.0f, r_max, inv, bins);
        if (bin < 0)
        {
            continue;
        }
        const double* x = values + 2 * (size_t) nl_ij[2 * k + 1];
        const double* y = query_values + 2 * (size_t) nl_ij[2 * k];
        double re = x[0] * y[0] + x[1] * y[1];
        double im = x[0] * y[1] - x[1] * y[0];
        counts[bin] += 1;
        correlation[2 * bin] += re;
        correlation[2 * bin + 1] += im;
    }
    for (uint32_t b = 0; b < bins; ++b)
    {
        if (counts[b] != 0)
        {
            correlation[2 * b] /= (double) counts[b];
            correlation[2 * b + 1] /= (double) counts[b];
        }
    }
    return 0;
}

/* ---- PMFTXY ------------------------------------------------------------------------------------------
 * PMFTXY::accumulate (freud/pmft/PMFTXY.cc:65-87) over the bonds of a NeighborList: the bond vector rotated by
 * rotmat2::fromAngle(-theta_i) (VectorMath.h:912-936: rows (cos, -sin), (sin, cos); products then one add per row),
 * binned on RegularAxis(n_x, -x_max, x_max) x RegularAxis(n_y, -y_max, y_max) with linear index bx * n_y + by
 * (Histogram.h:301-351; a bond outside either axis is dropped); then PMFT::reduce (freud/pmft/PMFT.h:73-83) with the
 * constant Jacobian factor 1 / (dx dy) (PMFTXY.cc:49-51, 59-63).  cosf / sinf are this machine's libm, as upstream. */
int fport_pmftxy(const uint32_t* nl_ij, const float* nl_v, uint64_t n_bonds, const float* query_orientations,
                 float x_max, float y_max, uint32_t n_x, uint32_t n_y, float box_volume, uint32_t n_points,
                 uint32_t n_query, uint32_t* counts, float* pcf)
{
    volatile float wx_v = (x_max - (-x_max)) / (float) n_x, wy_v = (y_max - (-y_max)) / (float) n_y;
    volatile float ix_v = 1.0f / wx_v, iy_v = 1.0f / wy_v;
    float inv_x = ix_v, inv_y = iy_v;
    memset(counts, 0, (size_t) n_x * n_y * sizeof(uint32_t));
    for (uint64_t k = 0; k < n_bonds; ++k)
    {
        float t = -query_orientations[nl_ij[2 * k]];
        float c = cosf(t), s = sinf(t);
        float vx = nl_v[3 * k], vy = nl_v[3 * k + 1];
        float ms = -s;
        float a1 = c * vx, a2 = ms * vy;
        float rx = a1 + a2;
        float b1 = s * vx, b2 = c * vy;
        float ry = b1 + b2;
        int64_t bx = axis_bin(rx, -x_max, x_max, inv_x, n_x);
        int64_t by = axis_bin(ry, -y_max, y_max, inv_y, n_y);
        if (bx >= 0 && by >= 0)
        {
            counts[(size_t) bx * n_y + (size_t) by] += 1;
        }
    }
    volatile float dx = 2.0f * x_max / (float) n_x, dy = 2.0f * y_max / (float) n_y;
    volatile float jac = dx * dy;
    volatile float inv_num_dens = box_volume / (float) n_query;
    volatile float den = 1.0f * (float) n_points; /* one frame */
    volatile float norm_factor = 1.0f / den;
    volatile float prefactor = inv_num_dens * norm_factor;
    volatile float jf = 1.0f / jac;
    for (size_t i = 0; i < (size_t) n_x * n_y; ++i)
    {
        volatile float t = (float) counts[i] * prefactor;
        pcf[i] = t * jf;
    }
    return 0;
}

/* ---- PMFTXYZ / PMFTXYT / PMFTR12 over the bonds of a NeighborList (one frame) ------------------------------------
 * kind 0 = XYZ (freud/pmft/PMFTXYZ.cc:24-147): v = rotate(equiv[e], rotate(conj(q_i), delta)) for every equivalent
 *          orientation e (VectorMath.h:765, 810-818), bins on x, y, z in [-max, max]; orientations are quaternions
 *          (s, x, y, z); reduce divides by n_equiv as well (PMFTXYZ.cc:86-98).
 * kind 1 = XYT (PMFTXYT.cc:28-101): (x, y) = rotmat2(-theta_i) * delta, t = modulusPositive(theta_j - atan2f(-dy, -dx),
 *          2 pi); Jacobian dx dy (1 / n_t).
 * kind 2 = R12 (PMFTR12.cc:28-113): r = bond distance, t1 = modulusPositive(theta_j - atan2f(dy, dx), 2 pi),
 *          t2 = modulusPositive(theta_i - atan2f(-dy, -dx), 2 pi); inverse Jacobian 1 / (r_centre dr (2 pi / n_t1) (1 / n_t2)).
 * cosf / sinf / atan2f / fmodf are this machine's libm, as upstream.  Linear bin index (b0 n1 + b1) n2 + b2
 * (Histogram.h:327-351); PMFT::reduce as in freud/pmft/PMFT.h:73-83. */
static void quat_rotate(float s, float qx, float qy, float qz, float* x, float* y, float* z)
{
    float bx = *x, by = *y, bz = *z;
    float p1 = qx * qx, p2 = qy * qy, p3 = qz * qz;
    float vv = (p1 + p2) + p3;
    float ss = s * s;
    float a = ss - vv;
    float two_s = 2.0f * s;
    float c1 = qy * bz, c2 = qz * by, c3 = qz * bx, c4 = qx * bz, c5 = qx * by, c6 = qy * bx;
    float cx = c1 - c2, cy = c3 - c4, cz = c5 - c6;
    float d1 = qx * bx, d2 = qy * by, d3 = qz * bz;
    float vb = (d1 + d2) + d3;
    float two_vb = 2.0f * vb;
    float t1x = bx * a, t2x = cx * two_s, t3x = qx * two_vb;
    float t1y = by * a, t2y = cy * two_s, t3y = qy * two_vb;
    float t1z = bz * a, t2z = cz * two_s, t3z = qz * two_vb;
    *x = (t1x + t2x) + t3x;
    *y = (t1y + t2y) + t3y;
    *z = (t1z + t2z) + t3z;
}

static float mod_two_pi(float a)
{
    const float two_pi = (float) (2.0 * M_PI); /* Box.h:24 */
    float inner = fmodf(a, two_pi) + two_pi;
    return fmodf(inner, two_pi);
}

int fport_pmft3(int kind, const uint32_t* nl_ij, const float* nl_v, const float* nl_d, uint64_t n_bonds,
                const float* orientations, const float* query_orientations, const float* equiv, uint32_t n_equiv,
                float max0, float max1, float max2, uint32_t n0, uint32_t n1, uint32_t n2, float box_volume,
                uint32_t n_points, uint32_t n_query, uint32_t* counts, float* pcf)
{
    const float two_pi = (float) (2.0 * M_PI);
    float lo[3], hi[3], inv[3], width[3];
    uint32_t n[3] = {n0, n1, n2};
    float mx[3] = {max0, max1, max2};
    for (int ax = 0; ax < 3; ++ax)
    {
        int const angle = (kind == 1 && ax == 2) || (kind == 2 && ax > 0);
        lo[ax] = angle || kind == 2 ? 0.0f : -mx[ax];
        hi[ax] = angle ? two_pi : mx[ax];
        axis_params(n[ax], lo[ax], hi[ax], &width[ax], &inv[ax]);
    }
    size_t const n_bins = (size_t) n0 * n1 * n2;
    memset(counts, 0, n_bins * sizeof(uint32_t));
    for (uint64_t k = 0; k < n_bonds; ++k)
    {
        uint32_t const i = nl_ij[2 * k], j = nl_ij[2 * k + 1];
        float const dx = nl_v[3 * k], dy = nl_v[3 * k + 1], dz = nl_v[3 * k + 2];
        if (kind == 0)
        {
            const float* q = query_orientations + 4 * (size_t) i;
            float x = dx, y = dy, z = dz;
            quat_rotate(q[0], -q[1], -q[2], -q[3], &x, &y, &z);
            for (uint32_t e = 0; e < n_equiv; ++e)
            {
                float ex = x, ey = y, ez = z;
                quat_rotate(equiv[4 * e], equiv[4 * e + 1], equiv[4 * e + 2], equiv[4 * e + 3], &ex, &ey, &ez);
                int64_t b0 = axis_bin(ex, lo[0], hi[0], inv[0], n0), b1 = axis_bin(ey, lo[1], hi[1], inv[1], n1),
                        b2 = axis_bin(ez, lo[2], hi[2], inv[2], n2);
                if (b0 >= 0 && b1 >= 0 && b2 >= 0)
                {
                    counts[((size_t) b0 * n1 + (size_t) b1) * n2 + (size_t) b2] += 1;
                }
            }
            continue;
        }
        int64_t b0, b1, b2;
        if (kind == 1)
        {
            float t = -query_orientations[i];
            float c = cosf(t), sn = sinf(t), ms = -sn;
            float a1 = c * dx, a2 = ms * dy, g1 = sn * dx, g2 = c * dy;
            float rx = a1 + a2, ry = g1 + g2;
            float d_theta = atan2f(-dy, -dx);
            float arg = orientations[j] - d_theta;
            b0 = axis_bin(rx, lo[0], hi[0], inv[0], n0);
            b1 = axis_bin(ry, lo[1], hi[1], inv[1], n1);
            b2 = axis_bin(mod_two_pi(arg), lo[2], hi[2], inv[2], n2);
        }
        else
        {
            float d_theta1 = atan2f(dy, dx), d_theta2 = atan2f(-dy, -dx);
            float arg1 = orientations[j] - d_theta1, arg2 = query_orientations[i] - d_theta2;
            b0 = axis_bin(nl_d[k], lo[0], hi[0], inv[0], n0);
            b1 = axis_bin(mod_two_pi(arg1), lo[1], hi[1], inv[1], n1);
            b2 = axis_bin(mod_two_pi(arg2), lo[2], hi[2], inv[2], n2);
        }
        if (b0 >= 0 && b1 >= 0 && b2 >= 0)
        {
            counts[((size_t) b0 * n1 + (size_t) b1) * n2 + (size_t) b2] += 1;
        }
    }
    float inv_num_dens = box_volume / (float) n_query;
    float den = 1.0f * (float) n_points; /* one frame */
    if (kind == 0)
    {
        den = den * (float) n_equiv;
    }
    float norm_factor = 1.0f / den;
    float prefactor = inv_num_dens * norm_factor;
    float jf = 0.0f;
    if (kind == 0)
    {
        float ddx = 2.0f * max0 / (float) n0, ddy = 2.0f * max1 / (float) n1, ddz = 2.0f * max2 / (float) n2;
        float jac = ddx * ddy;
        jac = jac * ddz;
        jf = 1.0f / jac;
    }
    else if (kind == 1)
    {
        float ddx = 2.0f * max0 / (float) n0, ddy = 2.0f * max1 / (float) n1, dt = 1 / (float) n2;
        float jac = ddx * ddy;
        jac = jac * dt;
        jf = 1.0f / jac;
    }
    float dr = max0 / (float) n0, dt1 = two_pi / (float) n1, dt2 = 1 / (float) n2;
    float product = dr * dt1;
    product = product * dt2;
    for (size_t b = 0; b < n_bins; ++b)
    {
        float f = jf;
        if (kind == 2)
        {
            size_t const ir = b / ((size_t) n1 * n2);
            float w = (float) ir * width[0], w1 = (float) (ir + 1) * width[0];
            float e0 = lo[0] + w, e1 = lo[0] + w1; /* RegularAxis edges, Histogram.h:133-137 */
            float sum = e0 + e1;
            float r = sum / 2.0f;
            float rp = r * product;
            f = 1.0f / rp;
        }
        float t = (float) counts[b] * prefactor;
        pcf[b] = t * f;
    }
    return 0;
}

/* ---- BondOrder over the bonds of a NeighborList (one frame), freud/environment/BondOrder.cc:30-153 ---------------
 * mode 0 bod: v = bond vector; 1 lbod: v = rotate(conj(o_j), v); 2 obcd: v = rotate(q_i, rotate(conj(o_j), v));
 * 3 oocd: v = rotate(conj(o_j), rotate(q_i, z)).  theta = modulusPositive(atan2f(v.y, v.x), 2 pi),
 * phi = acosf(v.z / sqrtf(v.v)); bins on RegularAxis(n_theta, 0, 2 pi) x RegularAxis(n_phi, 0, pi); the diagram divides
 * the counts by the solid angle of the bin, dt (cos(phi_j) - cos(phi_j + dp)), and by the number of frames (:88-93). */
int fport_bond_order(int mode, const uint32_t* nl_ij, const float* nl_v, uint64_t n_bonds, const float* orientations,
                     const float* query_orientations, uint32_t n_theta, uint32_t n_phi, uint32_t* counts, float* bond_order)
{
    const float two_pi = (float) (2.0 * M_PI), pi_f = (float) M_PI;
    float wt, it, wp, ip;
    axis_params(n_theta, 0.0f, two_pi, &wt, &it);
    axis_params(n_phi, 0.0f, pi_f, &wp, &ip);
    memset(counts, 0, (size_t) n_theta * n_phi * sizeof(uint32_t));
    for (uint64_t k = 0; k < n_bonds; ++k)
    {
        uint32_t const i = nl_ij[2 * k], j = nl_ij[2 * k + 1];
        float x = nl_v[3 * k], y = nl_v[3 * k + 1], z = nl_v[3 * k + 2];
        if (mode != 0)
        {
            const float* rq = orientations + 4 * (size_t) j;
            const float* q = query_orientations + 4 * (size_t) i;
            if (mode == 3)
            {
                x = 0.0f;
                y = 0.0f;
                z = 1.0f;
                quat_rotate(q[0], q[1], q[2], q[3], &x, &y, &z);
            }
            quat_rotate(rq[0], -rq[1], -rq[2], -rq[3], &x, &y, &z);
            if (mode == 2)
            {
                quat_rotate(q[0], q[1], q[2], q[3], &x, &y, &z);
            }
        }
        float theta = mod_two_pi(atan2f(y, x));
        float xx = x * x, yy = y * y, zz = z * z;
        float dot = (xx + yy) + zz;
        float arg = z / sqrtf(dot);
        float phi = acosf(arg);
        int64_t bt = axis_bin(theta, 0.0f, two_pi, it, n_theta), bp = axis_bin(phi, 0.0f, pi_f, ip, n_phi);
        if (bt >= 0 && bp >= 0)
        {
            counts[(size_t) bt * n_phi + (size_t) bp] += 1;
        }
    }
    float dt = two_pi / (float) n_theta;
    float dp = (float) (M_PI / (double) (float) n_phi);
    for (uint32_t a = 0; a < n_theta; ++a)
    {
        for (uint32_t b = 0; b < n_phi; ++b)
        {
            float phi = (float) b * dp;
            float phi2 = phi + dp;
            float diff = cosf(phi) - cosf(phi2);
            float sa = dt * diff;
            float t = (float) counts[(size_t) a * n_phi + b] / sa;
            bond_order[(size_t) a * n_phi + b] = t / 1.0f; /* one frame */
        }
    }
    return 0;
}

