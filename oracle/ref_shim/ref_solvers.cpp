/* C harness over the reference's header-only solvers. TEST INFRASTRUCTURE ONLY.
 *
 * Compiles algebra/sfem_cg.hpp, sfem_Chebyshev3.hpp, sfem_PowerMethod.hpp, sfem_Stationary.hpp,
 * sfem_ShiftableJacobi.hpp and sfem_Multigrid.hpp WHERE THEY LIE under /root/reference (oracle/Makefile target
 * `ref`), against the smesh stand-in of oracle/ref_shim/sfem_aliases.hpp. Nothing of the reference is copied:
 * this file only instantiates its templates and exposes them as an object graph with C linkage so that
 * tests/ can
 *   (1) run the reference's own CG / PCG / Chebyshev / Jacobi / multigrid on the CPU (OpenMP_BLAS + an operator
 *       callback that calls the reference's element kernels) and read iteration counts and solutions: the oracle
 *       for "CG/MG iteration counts must be identical";
 *   (2) run the SAME reference solver objects with execution_space = DEVICE, a BLAS<T> whose 13 virtuals are bound
 *       to the b200fem C ABI (include/b200fem.h) and an operator callback that calls b200fem_op_apply on device
 *       pointers: the literal drop-in of BASELINE.json's north_star ("sfem_cg.hpp and sfem_Multigrid.hpp use it
 *       as a drop-in").
 * The library links nothing of sfem_b200/: the device entry points arrive as function pointers. */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "sfem_base.hpp"

#include "sfem_aliases.hpp"

#include "sfem_Chebyshev3.hpp"
#include "sfem_MixedPrecisionShiftableBlockSymJacobi.hpp"
#include "sfem_Multigrid.hpp"
#include "sfem_ShiftableJacobi.hpp"
#include "sfem_Stationary.hpp"
#include "sfem_cg.hpp"

namespace smesh {
    extern "C" {
    void *(*oracle_shim_device_calloc)(size_t bytes) = nullptr;
    void (*oracle_shim_device_free)(void *ptr)       = nullptr;
    }
}  // namespace smesh

using namespace sfem;

namespace {

    /* b200fem C-ABI signatures (include/b200fem.h), bound at run time */
    struct DeviceTable {
        int (*malloc_)(void **, size_t);
        int (*free_)(void *);
        int (*zeros)(int, ptrdiff_t, void *, void *);
        int (*values)(int, ptrdiff_t, double, void *, void *);
        int (*copy)(int, ptrdiff_t, const void *, void *, void *);
        int (*scal)(int, ptrdiff_t, double, void *, void *);
        int (*axpy)(int, ptrdiff_t, double, const void *, void *, void *);
        int (*axpby)(int, ptrdiff_t, double, const void *, double, void *, void *);
        int (*zaxpby)(int, ptrdiff_t, double, const void *, double, const void *, void *, void *);
        int (*xypaz)(int, ptrdiff_t, const void *, const void *, double, void *, void *);
        int (*reciprocal)(int, ptrdiff_t, double, void *, void *);
        int (*dot)(int, ptrdiff_t, const void *, const void *, double *, void *);
        int (*norm2)(int, ptrdiff_t, const void *, double *, void *);
    };
    DeviceTable g_dev{};
    bool        g_dev_set = false;

    void *dev_calloc(size_t bytes) {
        void *p = nullptr;
        if (g_dev.malloc_(&p, bytes)) SFEM_ERROR("device allocation failed\n");
        return p;
    }
    void dev_free(void *p) { g_dev.free_(p); }

    constexpr int RT = (int)sizeof(real_t);

    /* sfem::BLAS<real_t> over the b200fem vector entry points */
    class TableBLAS final : public BLAS<real_t> {
    public:
        real_t *allocate(const std::size_t n) override { return (real_t *)dev_calloc((n ? n : 1) * sizeof(real_t)); }
        void    destroy(void *a) override { dev_free(a); }
        void    zeros(const std::size_t n, real_t *const x) override { g_dev.zeros(RT, (ptrdiff_t)n, x, nullptr); }
        void    values(const std::size_t n, const real_t v, real_t *const x) override {
            g_dev.values(RT, (ptrdiff_t)n, (double)v, x, nullptr);
        }
        void copy(const ptrdiff_t n, const real_t *const s, real_t *const d) override { g_dev.copy(RT, n, s, d, nullptr); }
        real_t dot(const ptrdiff_t n, const real_t *const l, const real_t *const r) override {
            double v = 0;
            g_dev.dot(RT, n, l, r, &v, nullptr);
            return (real_t)v;
        }
        void axpy(const ptrdiff_t n, const real_t a, const real_t *const x, real_t *const y) override {
            g_dev.axpy(RT, n, (double)a, x, y, nullptr);
        }
        void axpby(const ptrdiff_t n, const real_t a, const real_t *const x, const real_t b, real_t *const y) override {
            g_dev.axpby(RT, n, (double)a, x, (double)b, y, nullptr);
        }
        void scal(const std::ptrdiff_t n, const real_t a, real_t *const x) override { g_dev.scal(RT, n, (double)a, x, nullptr); }
        void reciprocal(const std::ptrdiff_t n, const real_t a, real_t *const x) override {
            g_dev.reciprocal(RT, n, (double)a, x, nullptr);
        }
        real_t norm2(const ptrdiff_t n, const real_t *const x) override {
            double v = 0;
            g_dev.norm2(RT, n, x, &v, nullptr);
            return (real_t)v;
        }
        void zaxpby(const ptrdiff_t n, const real_t a, const real_t *const x, const real_t b, const real_t *const y,
                    real_t *const z) override {
            g_dev.zaxpby(RT, n, (double)a, x, (double)b, y, z, nullptr);
        }
        void xypaz(const ptrdiff_t n, const real_t *const x, const real_t *const y, const real_t a,
                   real_t *const z) override {
            g_dev.xypaz(RT, n, x, y, (double)a, z, nullptr);
        }
    };

    std::shared_ptr<BLAS<real_t>> blas_for(int es) {
        if (es == EXECUTION_SPACE_DEVICE) {
            if (!g_dev_set) SFEM_ERROR("refsolv: device BLAS table not registered\n");
            return std::make_shared<TableBLAS>();
        }
        return make_openmp_blas<real_t>();
    }

    struct Obj {
        std::shared_ptr<Operator<real_t>> op;
        std::vector<std::shared_ptr<void>> keep; /* anything the op references */
    };

    typedef int (*apply_fn)(void *ctx, const real_t *x, real_t *y);

}  // namespace

extern "C" {

int refsolv_sizeof_real() { return RT; }

/* fns: malloc, free, zeros, values, copy, scal, axpy, axpby, zaxpby, xypaz, reciprocal, dot, norm2 */
int refsolv_set_device_table(void **fns) {
    static_assert(sizeof(DeviceTable) == 13 * sizeof(void *), "table layout");
    memcpy(&g_dev, fns, sizeof(DeviceTable));
    g_dev_set                        = true;
    smesh::oracle_shim_device_calloc = dev_calloc;
    smesh::oracle_shim_device_free   = dev_free;
    return 0;
}

void *refsolv_make_op(ptrdiff_t rows, ptrdiff_t cols, apply_fn fn, void *ctx, int es) {
    auto o = new Obj();
    o->op  = make_op<real_t>(
            rows, cols, [fn, ctx](const real_t *const x, real_t *const y) { fn(ctx, x, y); }, (ExecutionSpace)es);
    return o;
}

void *refsolv_cg(void *op, void *precond, double rtol, double atol, int max_it, int es) {
    auto cg = std::make_shared<ConjugateGradient<real_t>>();
    cg->blas             = blas_for(es);
    cg->execution_space_ = (ExecutionSpace)es;
    cg->verbose          = false;
    cg->set_op(((Obj *)op)->op);
    if (precond) cg->set_preconditioner_op(((Obj *)precond)->op);
    cg->set_rtol((real_t)rtol);
    cg->set_atol((real_t)atol);
    cg->set_max_it(max_it);
    auto o = new Obj();
    o->op  = cg;
    return o;
}

/* verbose = the reference's own monitor() lines on stdout ("<it>: residual abs: ..., rel: ..."), sfem_cg.hpp:184-199 */
void refsolv_cg_set_verbose(void *cg, int val) {
    std::static_pointer_cast<ConjugateGradient<real_t>>(((Obj *)cg)->op)->verbose = val != 0;
}

void refsolv_cg_set_host_fusion(void *cg, int val) {
    std::static_pointer_cast<ConjugateGradient<real_t>>(((Obj *)cg)->op)->set_host_fusion(val != 0);
}

/* diag: caller-owned (host or device according to es), wrapped without copy; must outlive the object */
void *refsolv_jacobi(real_t *diag, ptrdiff_t n, double relaxation, int es) {
    auto j                  = std::make_shared<ShiftableJacobi<real_t>>();
    j->blas                 = blas_for(es);
    j->execution_space_     = (ExecutionSpace)es;
    j->relaxation_parameter = (real_t)relaxation;
    j->set_diag(Buffer<real_t>::wrap(n, diag, (MemorySpace)es));
    auto o = new Obj();
    o->op  = j;
    return o;
}

/* host only: the reference's block smoother kernels are OpenMP code */
void *refsolv_block_sym_jacobi(real_t *diag6, ptrdiff_t n_blocks, mask_t *mask, ptrdiff_t mask_bytes,
                               double relaxation) {
    auto j = std::make_shared<ShiftableBlockSymJacobi<real_t>>();
    j->default_init();
    j->relaxation_parameter = (real_t)relaxation;
    j->constraints_mask     = Buffer<mask_t>::wrap(mask_bytes, mask);
    j->set_diag(Buffer<real_t>::wrap(n_blocks * 6, diag6));
    auto o = new Obj();
    o->op  = j;
    return o;
}

#ifndef ORACLE_REAL_FLOAT32
/* MixedPrecisionShiftableBlockSymJacobi<double, float> (algebra/sfem_MixedPrecisionShiftableBlockSymJacobi.hpp), host only */
void *refsolv_mixed_block_sym_jacobi(real_t *diag6, ptrdiff_t n_blocks, mask_t *mask, ptrdiff_t mask_bytes, double relaxation) {
    auto j = std::make_shared<MixedPrecisionShiftableBlockSymJacobi<double, float>>();
    j->default_init();
    j->relaxation_parameter = relaxation;
    j->constraints_mask     = Buffer<mask_t>::wrap(mask_bytes, mask);
    j->set_diag(Buffer<double>::wrap(n_blocks * 6, diag6));
    auto o = new Obj();
    o->op  = j;
    return o;
}
#endif

void *refsolv_stationary(void *op, void *precond, int max_it, int es) {
    auto s              = std::make_shared<StationaryIteration<real_t>>();
    s->op               = ((Obj *)op)->op;
    s->preconditioner   = ((Obj *)precond)->op;
    s->n_dofs           = s->op->cols();
    s->blas             = blas_for(es);
    s->execution_space_ = (ExecutionSpace)es;
    s->set_max_it(max_it);
    auto o = new Obj();
    o->op  = s;
    return o;
}

void *refsolv_cheb3(void *op, double eig_tol, int eig_max_it, double scale_eig_max, double scale_eig_min, int max_it,
                    int initial_guess_zero, int es) {
    auto c              = std::make_shared<Chebyshev3<real_t>>();
    c->n_dofs           = ((Obj *)op)->op->rows();
    c->set_op(((Obj *)op)->op);
    c->blas             = blas_for(es);
    c->execution_space_ = (ExecutionSpace)es;
    c->ensure_power_method();
    c->eigen_solver_tol    = (real_t)eig_tol;
    c->eigen_solver_max_it = eig_max_it;
    c->scale_eig_max       = (real_t)scale_eig_max;
    c->scale_eig_min       = (real_t)scale_eig_min;
    c->set_max_it(max_it);
    c->set_initial_guess_zero(initial_guess_zero != 0);
    auto o = new Obj();
    o->op  = c;
    return o;
}

void refsolv_cheb3_init_with_ones(void *c) { std::static_pointer_cast<Chebyshev3<real_t>>(((Obj *)c)->op)->init_with_ones(); }
void refsolv_cheb3_init(void *c, const real_t *guess) { std::static_pointer_cast<Chebyshev3<real_t>>(((Obj *)c)->op)->init(guess); }
double refsolv_cheb3_eig_max(void *c) { return (double)std::static_pointer_cast<Chebyshev3<real_t>>(((Obj *)c)->op)->eig_max; }
void refsolv_cheb3_set_initial_guess_zero(void *c, int v) {
    std::static_pointer_cast<Chebyshev3<real_t>>(((Obj *)c)->op)->set_initial_guess_zero(v != 0);
}

void *refsolv_mg(int es, int max_it, double atol, int cycle_type, int verbose) {
    auto mg = std::make_shared<Multigrid<real_t>>();
    mg->blas() = blas_for(es);
    mg->set_execution_space((ExecutionSpace)es);
    mg->set_max_it(max_it);
    mg->set_atol((real_t)atol);
    mg->set_cycle_type(cycle_type);
    mg->verbose = verbose != 0;
    auto o      = new Obj();
    o->op       = mg;
    return o;
}

void refsolv_mg_add_level(void *mg, void *op, void *smoother, void *prolongation, void *restriction) {
    auto m = std::static_pointer_cast<Multigrid<real_t>>(((Obj *)mg)->op);
    m->add_level(op ? ((Obj *)op)->op : nullptr, ((Obj *)smoother)->op, prolongation ? ((Obj *)prolongation)->op : nullptr,
                 restriction ? ((Obj *)restriction)->op : nullptr);
}

int refsolv_apply(void *obj, const real_t *b, real_t *x) { return ((Obj *)obj)->op->apply(b, x); }

int refsolv_iterations(void *obj) {
    auto s = std::dynamic_pointer_cast<MatrixFreeLinearSolver<real_t>>(((Obj *)obj)->op);
    return s ? s->iterations() : -1;
}

void refsolv_destroy(void *obj) { delete (Obj *)obj; }

/* PowerMethod on its own (algebra/sfem_PowerMethod.hpp:18-55), host BLAS */
double refsolv_power_method(void *op, int max_it, double tol, ptrdiff_t n, real_t *eigen_vector, real_t *work, int es) {
    auto            blas = blas_for(es);
    PowerMethod<real_t> pm;
    pm.norm2 = [blas](const std::ptrdiff_t n, const real_t *const x) { return blas->norm2(n, x); };
    pm.scal  = [blas](const std::ptrdiff_t n, const real_t a, real_t *const x) { blas->scal(n, a, x); };
    pm.zeros = [blas](const std::size_t n, real_t *const x) { blas->zeros(n, x); };
    auto o   = ((Obj *)op)->op;
    return (double)pm.max_eigen_value([o](const real_t *const x, real_t *const y) { o->apply(x, y); }, max_it, (real_t)tol, n,
                                      eigen_vector, work);
}
}
