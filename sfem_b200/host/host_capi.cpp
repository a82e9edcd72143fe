/* C entry points over the C++ host layer (b200fem_host.hpp), for tests/, bench.py and embedding from other
 * languages. Objects form a graph of boxed std::shared_ptr<Operator<T>>; every constructor takes `real_type`
 * (B200FEM_F64 / B200FEM_F32). Vectors are device pointers. Built as sfem_b200/libb200fem_host.so, linked against
 * libb200fem.so. The same object-graph shape is exposed for the reference's own solver headers by
 * oracle/ref_shim/ref_solvers.cpp so that parity tests read side by side. */
#include <cstring>

#include "b200fem_host.hpp"
#include "b200fem_mesh_io.hpp"

using namespace b200fem;

namespace {

    struct Box {
        int                   real_type{0};
        std::shared_ptr<void> obj;       /* shared_ptr<Operator<T>> target */
        std::shared_ptr<void> function;  /* shared_ptr<Function<T>> when the box is a function */
        std::shared_ptr<void> blas;      /* shared_ptr<DeviceBLAS<T>> */
        int                   kind{0};
    };
    enum { K_OP = 1, K_CG, K_JACOBI, K_BJACOBI, K_STATIONARY, K_CHEB, K_MG, K_FUNCTION, K_BLAS, K_COMM, K_HALO };

    /* distributed objects are not templated on the scalar type: boxed through `obj` as shared_ptr<void> */
    std::shared_ptr<Communicator> comm_of(void *b) {
        return (b && ((Box *)b)->kind == K_COMM) ? std::static_pointer_cast<Communicator>(((Box *)b)->obj) : nullptr;
    }
    std::shared_ptr<HaloExchange> halo_of(void *b) {
        return (b && ((Box *)b)->kind == K_HALO) ? std::static_pointer_cast<HaloExchange>(((Box *)b)->obj) : nullptr;
    }

    template <typename T>
    std::shared_ptr<Operator<T>> op_of(void *b) {
        return b ? std::static_pointer_cast<Operator<T>>(((Box *)b)->obj) : nullptr;
    }
    template <typename T>
    std::shared_ptr<DeviceBLAS<T>> blas_of(void *b) {
        return std::static_pointer_cast<DeviceBLAS<T>>(((Box *)b)->blas);
    }
    template <typename T, typename S>
    Box *box(const std::shared_ptr<S> &p, int kind, void *blas_box) {
        auto b       = new Box();
        b->real_type = RealType<T>::value;
        b->obj       = std::static_pointer_cast<Operator<T>>(p);
        b->kind      = kind;
        if (blas_box) b->blas = ((Box *)blas_box)->blas;
        return b;
    }

#define DISPATCH(rt, expr_double, expr_float) ((rt) == B200FEM_F64 ? (expr_double) : (expr_float))

    typedef int (*apply_cb)(void *ctx, const void *x, void *y);
    typedef double (*reduce_cb)(void *ctx, double v);

    template <typename T>
    Box *make_blas(reduce_cb cb, void *ctx) {
        auto bl = make_device_blas<T>();
        if (cb) bl->allreduce_sum = [cb, ctx](double v) { return cb(ctx, v); };
        auto b       = new Box();
        b->real_type = RealType<T>::value;
        b->blas      = bl;
        b->kind      = K_BLAS;
        return b;
    }

    template <typename T>
    Box *make_function(ptrdiff_t n_nodes, int block_size) {
        auto f       = std::make_shared<Function<T>>(n_nodes, block_size);
        auto b       = new Box();
        b->real_type = RealType<T>::value;
        b->function  = f;
        b->obj       = f->linear_op(f);
        b->kind      = K_FUNCTION;
        return b;
    }
    template <typename T>
    std::shared_ptr<Function<T>> fn_of(void *b) {
        return b ? std::static_pointer_cast<Function<T>>(((Box *)b)->function) : nullptr;
    }

    template <typename T>
    Box *make_callback_op(ptrdiff_t rows, ptrdiff_t cols, apply_cb fn, apply_cb overwrite, void *ctx) {
        typename LambdaOperator<T>::Fn a = [fn, ctx](const T *const x, T *const y) { return fn(ctx, x, y); };
        typename LambdaOperator<T>::Fn o = nullptr;
        if (overwrite) o = [overwrite, ctx](const T *const x, T *const y) { return overwrite(ctx, x, y); };
        return box<T>(make_op<T>(rows, cols, a, o), K_OP, nullptr);
    }

    template <typename T>
    Box *make_cg(void *blas, void *op, void *precond, double rtol, double atol, int max_it) {
        auto cg  = std::make_shared<ConjugateGradient<T>>();
        cg->blas = blas_of<T>(blas);
        cg->set_op(op_of<T>(op));
        if (precond) cg->set_preconditioner_op(op_of<T>(precond));
        cg->set_rtol((T)rtol);
        cg->set_atol((T)atol);
        cg->set_max_it(max_it);
        return box<T>(cg, K_CG, blas);
    }
    template <typename T>
    Box *make_jacobi(void *blas, const void *diag, ptrdiff_t n, double relaxation) {
        auto j                  = std::make_shared<ShiftableJacobi<T>>();
        j->blas                 = blas_of<T>(blas);
        j->relaxation_parameter = (T)relaxation;
        j->set_diag((const T *)diag, n);
        return box<T>(j, K_JACOBI, blas);
    }
    template <typename T>
    Box *make_bjacobi(void *blas, const void *diag6, ptrdiff_t n_blocks, const void *mask_dev, double relaxation) {
        auto j                  = std::make_shared<ShiftableBlockSymJacobi<T>>();
        j->blas                 = blas_of<T>(blas);
        j->relaxation_parameter = (T)relaxation;
        j->constraints_mask     = mask_dev;
        if (j->set_diag((const T *)diag6, n_blocks)) return nullptr;
        return box<T>(j, K_BJACOBI, blas);
    }
    template <typename T>
    Box *make_stationary(void *blas, void *op, void *precond, int max_it) {
        auto s            = std::make_shared<StationaryIteration<T>>();
        s->blas           = blas_of<T>(blas);
        s->op             = op_of<T>(op);
        s->preconditioner = op_of<T>(precond);
        s->n_dofs         = s->op->cols();
        s->set_max_it(max_it);
        return box<T>(s, K_STATIONARY, blas);
    }
    template <typename T>
    Box *make_cheb(void *blas, void *op, double eig_tol, int eig_max_it, double scale_eig_max, double scale_eig_min, int max_it,
                   int initial_guess_zero) {
        auto c  = std::make_shared<Chebyshev3<T>>();
        c->blas = blas_of<T>(blas);
        c->set_op(op_of<T>(op));
        c->eigen_solver_tol    = (T)eig_tol;
        c->eigen_solver_max_it = eig_max_it;
        c->scale_eig_max       = (T)scale_eig_max;
        c->scale_eig_min       = (T)scale_eig_min;
        c->set_max_it(max_it);
        c->set_initial_guess_zero(initial_guess_zero != 0);
        return box<T>(c, K_CHEB, blas);
    }
    template <typename T>
    Box *make_mg(void *blas, int max_it, double atol, int cycle_type, int verbose) {
        auto m  = std::make_shared<Multigrid<T>>();
        m->blas = blas_of<T>(blas);
        m->set_max_it(max_it);
        m->set_atol((T)atol);
        m->set_cycle_type(cycle_type);
        m->verbose = verbose != 0;
        return box<T>(m, K_MG, blas);
    }
    template <typename T>
    int copy_history(const std::vector<double> &h, double *out, int cap) {
        const int n = (int)h.size();
        for (int i = 0; i < n && i < cap; i++) out[i] = h[i];
        return n;
    }
}  // namespace

extern "C" {

/* BLAS context: `allreduce` (may be NULL) is called with every locally reduced scalar and returns the global sum */
void *b200host_blas(int real_type, reduce_cb allreduce, void *ctx) {
    return DISPATCH(real_type, make_blas<double>(allreduce, ctx), make_blas<float>(allreduce, ctx));
}
/* custom allocator for the vectors the solvers allocate (see DeviceBLAS::alloc_hook) */
typedef void *(*alloc_cb)(void *ctx, size_t bytes);
typedef void (*free_cb)(void *ctx, void *ptr);
int b200host_blas_set_allocator(void *blas, alloc_cb alloc, free_cb release, void *ctx) {
    Box *b = (Box *)blas;
    if (!b || b->kind != K_BLAS) return 1;
    auto a = [alloc, ctx](std::size_t bytes) { return alloc(ctx, bytes); };
    auto f = [release, ctx](void *p) { release(ctx, p); };
    if (b->real_type == B200FEM_F64) {
        blas_of<double>(b)->alloc_hook = a;
        blas_of<double>(b)->free_hook  = f;
    } else {
        blas_of<float>(b)->alloc_hook = a;
        blas_of<float>(b)->free_hook  = f;
    }
    return 0;
}
long b200host_blas_launches(void *blas) {
    Box *b = (Box *)blas;
    return DISPATCH(b->real_type, blas_of<double>(b)->n_launches, blas_of<float>(b)->n_launches);
}

/* Function (frontend/sfem_Function.cpp): operators are b200fem_op_t plans (not owned), Dirichlet node lists are
 * DEVICE int32 arrays (not owned). The box doubles as the Operator of make_linear_op(f). */
void *b200host_function(int real_type, ptrdiff_t n_nodes, int block_size) {
    return DISPATCH(real_type, make_function<double>(n_nodes, block_size), make_function<float>(n_nodes, block_size));
}
int b200host_function_add_operator(void *f, b200fem_op_t op) {
    Box *b = (Box *)f;
    if (b->real_type == B200FEM_F64)
        fn_of<double>(b)->add_operator(op);
    else
        fn_of<float>(b)->add_operator(op);
    return 0;
}
int b200host_function_add_dirichlet(void *f, const int32_t *idx_dev, ptrdiff_t n, int component, double value) {
    Box               *b = (Box *)f;
    DirichletCondition c{idx_dev, n, component, value};
    if (b->real_type == B200FEM_F64)
        fn_of<double>(b)->add_constraint(c);
    else
        fn_of<float>(b)->add_constraint(c);
    return 0;
}
int b200host_function_apply_constraints(void *f, void *x) {
    Box *b = (Box *)f;
    return DISPATCH(b->real_type, fn_of<double>(b)->apply_constraints((double *)x), fn_of<float>(b)->apply_constraints((float *)x));
}
int b200host_function_apply_zero_constraints(void *f, void *x) {
    Box *b = (Box *)f;
    return DISPATCH(b->real_type, fn_of<double>(b)->apply_zero_constraints((double *)x),
                    fn_of<float>(b)->apply_zero_constraints((float *)x));
}
int b200host_function_set_value_to_constrained_dofs(void *f, double v, void *x) {
    Box *b = (Box *)f;
    return DISPATCH(b->real_type, fn_of<double>(b)->set_value_to_constrained_dofs(v, (double *)x),
                    fn_of<float>(b)->set_value_to_constrained_dofs((float)v, (float *)x));
}
int b200host_function_hessian_diag(void *f, void *d) {
    Box *b = (Box *)f;
    return DISPATCH(b->real_type, fn_of<double>(b)->hessian_diag(nullptr, (double *)d),
                    fn_of<float>(b)->hessian_diag(nullptr, (float *)d));
}
int b200host_function_hessian_block_diag_sym(void *f, void *d) {
    Box *b = (Box *)f;
    return DISPATCH(b->real_type, fn_of<double>(b)->hessian_block_diag_sym(nullptr, (double *)d),
                    fn_of<float>(b)->hessian_block_diag_sym(nullptr, (float *)d));
}
int b200host_function_launches_per_apply(void *f, int overwrite) {
    Box *b = (Box *)f;
    return DISPATCH(b->real_type, fn_of<double>(b)->launches_per_apply(overwrite != 0),
                    fn_of<float>(b)->launches_per_apply(overwrite != 0));
}

/* Operator from callbacks (used for the distributed operator, whose halo exchange lives in torch.distributed) */
void *b200host_callback_op(int real_type, ptrdiff_t rows, ptrdiff_t cols, apply_cb apply, apply_cb overwrite, void *ctx) {
    return DISPATCH(real_type, make_callback_op<double>(rows, cols, apply, overwrite, ctx),
                    make_callback_op<float>(rows, cols, apply, overwrite, ctx));
}

/* ---- distributed layer (Communicator, HaloExchange, ParallelMatrixFreeOperator, parallel CG) ---------------------- */
int b200host_comm_unique_id(void *id_out) { return b200fem_comm_unique_id(id_out); }
void *b200host_comm(int world, int rank, const void *unique_id) {
    auto c = Communicator::create(world, rank, unique_id);
    if (!c) return nullptr;
    auto b  = new Box();
    b->obj  = c;
    b->kind = K_COMM;
    return b;
}
int    b200host_comm_rank(void *c) { return comm_of(c)->rank(); }
int    b200host_comm_size(void *c) { return comm_of(c)->size(); }
double b200host_comm_sum(void *c, double v) { return comm_of(c)->sum(v); }
double b200host_comm_max(void *c, double v) { return comm_of(c)->max(v); }
int    b200host_comm_error(void *c) { return b200fem_comm_error(comm_of(c)->get()); }
void   b200host_comm_barrier(void *c) { comm_of(c)->barrier(); }

/* halo plan of one space; arguments as b200fem_halo_create (collective) */
void *b200host_halo(void *comm, int real_type, int block_size, ptrdiff_t n_owned, ptrdiff_t n_local, int n_send_peers,
                    const int *send_peers, const ptrdiff_t *send_offsets, const int32_t *send_idx, int n_recv_peers,
                    const int *recv_peers, const ptrdiff_t *recv_begin, const ptrdiff_t *recv_count, int transport) {
    auto c = comm_of(comm);
    if (!c) return nullptr;
    b200fem_halo_t h = nullptr;
    if (b200fem_halo_create(&h, c->get(), real_type, block_size, n_owned, n_local, n_send_peers, send_peers, send_offsets, send_idx,
                            n_recv_peers, recv_peers, recv_begin, recv_count, transport))
        return nullptr;
    auto b       = new Box();
    b->real_type = real_type;
    b->obj       = std::make_shared<HaloExchange>(h, c);
    b->kind      = K_HALO;
    return b;
}
int b200host_halo_set_recv_index(void *halo, const ptrdiff_t *recv_offsets, const int32_t *recv_idx) {
    return b200fem_halo_set_recv_index(halo_of(halo)->get(), recv_offsets, recv_idx);
}
int b200host_halo_gather(void *halo, void *x_local) { return halo_of(halo)->gather(x_local); }
int b200host_halo_transport(void *halo) { return b200fem_halo_transport(halo_of(halo)->get()); }
ptrdiff_t b200host_halo_values_sent(void *halo) { return b200fem_halo_values_sent(halo_of(halo)->get()); }
int b200host_halo_launches_per_gather(void *halo) { return b200fem_halo_launches_per_gather(halo_of(halo)->get()); }

/* ParallelMatrixFreeOperator over any local operator box (Function, restriction, prolongation): rows / cols = OWNED
 * dofs of the output / input space, `in_halo` exchanges the input space */
void *b200host_parallel_op(void *local_op, void *in_halo, ptrdiff_t owned_rows, ptrdiff_t owned_cols) {
    Box *o = (Box *)local_op;
    auto h = halo_of(in_halo);
    if (!o || !h) return nullptr;
    if (o->real_type == B200FEM_F64)
        return box<double>(std::make_shared<ParallelMatrixFreeOperator<double>>(op_of<double>(o), h, owned_rows, owned_cols), K_OP, nullptr);
    return box<float>(std::make_shared<ParallelMatrixFreeOperator<float>>(op_of<float>(o), h, owned_rows, owned_cols), K_OP, nullptr);
}
/* elements [e0, e1) of the local operator read no ghost value: run them between the two halves of the gather */
int b200host_parallel_op_set_independent_range(void *pop, ptrdiff_t e0, ptrdiff_t e1) {
    Box *o = (Box *)pop;
    if (!o || e0 < 0 || e1 < e0) return 1;
    if (o->real_type == B200FEM_F64) {
        auto p = std::dynamic_pointer_cast<ParallelMatrixFreeOperator<double>>(op_of<double>(o));
        if (!p) return 1;
        p->set_independent_range(e0, e1);
        return p->overlapped() ? 0 : 2;
    }
    auto p = std::dynamic_pointer_cast<ParallelMatrixFreeOperator<float>>(op_of<float>(o));
    if (!p) return 1;
    p->set_independent_range(e0, e1);
    return p->overlapped() ? 0 : 2;
}
/* reductions of this BLAS run over the communicator (device-side sum + ncclAllReduce + one host read) */
int b200host_blas_set_comm(void *blas, void *comm) {
    Box *b = (Box *)blas;
    auto c = comm_of(comm);
    if (!b || b->kind != K_BLAS || !c) return 1;
    if (b->real_type == B200FEM_F64)
        blas_of<double>(b)->comm = c->get();
    else
        blas_of<float>(b)->comm = c->get();
    b->function = c; /* keeps the communicator alive as long as the BLAS */
    return 0;
}
/* vectors of n_owned entries are allocated with n_local entries (ParallelOperator::row_allocation_size) */
int b200host_blas_reserve_allocation(void *blas, ptrdiff_t n_owned, ptrdiff_t n_local) {
    Box *b = (Box *)blas;
    if (!b || b->kind != K_BLAS) return 1;
    if (b->real_type == B200FEM_F64)
        blas_of<double>(b)->reserve_allocation((size_t)n_owned, (size_t)n_local);
    else
        blas_of<float>(b)->reserve_allocation((size_t)n_owned, (size_t)n_local);
    return 0;
}
long b200host_blas_reductions(void *blas) {
    Box *b = (Box *)blas;
    return DISPATCH(b->real_type, blas_of<double>(b)->n_reductions, blas_of<float>(b)->n_reductions);
}
/* BLAS calls on raw device vectors through a BLAS box (drivers: distributed norms / dots of their own vectors) */
double b200host_blas_dot(void *blas, ptrdiff_t n, const void *x, const void *y) {
    Box *b = (Box *)blas;
    return DISPATCH(b->real_type, (double)blas_of<double>(b)->dot(n, (const double *)x, (const double *)y),
                    (double)blas_of<float>(b)->dot(n, (const float *)x, (const float *)y));
}

/* level transfer operators; `constraints_function` (may be NULL) zeroes the constrained dofs of the result */
void *b200host_restriction(int real_type, b200fem_transfer_t t, int vec_size, void *coarse_function) {
    if (real_type == B200FEM_F64) return box<double>(make_restriction<double>(t, vec_size, fn_of<double>(coarse_function)), K_OP, nullptr);
    return box<float>(make_restriction<float>(t, vec_size, fn_of<float>(coarse_function)), K_OP, nullptr);
}
void *b200host_prolongation(int real_type, b200fem_transfer_t t, int vec_size, void *fine_function) {
    if (real_type == B200FEM_F64) return box<double>(make_prolongation<double>(t, vec_size, fn_of<double>(fine_function)), K_OP, nullptr);
    return box<float>(make_prolongation<float>(t, vec_size, fn_of<float>(fine_function)), K_OP, nullptr);
}

void *b200host_cg(void *blas, void *op, void *precond, double rtol, double atol, int max_it) {
    return DISPATCH(((Box *)blas)->real_type, make_cg<double>(blas, op, precond, rtol, atol, max_it),
                    make_cg<float>(blas, op, precond, rtol, atol, max_it));
}
void *b200host_jacobi(void *blas, const void *diag_dev, ptrdiff_t n, double relaxation) {
    return DISPATCH(((Box *)blas)->real_type, make_jacobi<double>(blas, diag_dev, n, relaxation),
                    make_jacobi<float>(blas, diag_dev, n, relaxation));
}
void *b200host_block_sym_jacobi(void *blas, const void *diag6_dev, ptrdiff_t n_blocks, const void *mask_dev, double relaxation) {
    return DISPATCH(((Box *)blas)->real_type, make_bjacobi<double>(blas, diag6_dev, n_blocks, mask_dev, relaxation),
                    make_bjacobi<float>(blas, diag6_dev, n_blocks, mask_dev, relaxation));
}
/* MixedPrecisionShiftableBlockSymJacobi<double, float>: fp64 only */
void *b200host_mixed_block_sym_jacobi(const void *diag6_dev, ptrdiff_t n_blocks, const void *mask_dev, double relaxation) {
    auto j                  = std::make_shared<MixedPrecisionShiftableBlockSymJacobi>();
    j->relaxation_parameter = relaxation;
    j->constraints_mask     = mask_dev;
    if (j->set_diag((const double *)diag6_dev, n_blocks)) return nullptr;
    return box<double>(j, K_BJACOBI, nullptr);
}
void *b200host_stationary(void *blas, void *op, void *precond, int max_it) {
    return DISPATCH(((Box *)blas)->real_type, make_stationary<double>(blas, op, precond, max_it),
                    make_stationary<float>(blas, op, precond, max_it));
}
void *b200host_cheb3(void *blas, void *op, double eig_tol, int eig_max_it, double scale_eig_max, double scale_eig_min,
                     int max_it, int initial_guess_zero) {
    return DISPATCH(((Box *)blas)->real_type,
                    make_cheb<double>(blas, op, eig_tol, eig_max_it, scale_eig_max, scale_eig_min, max_it, initial_guess_zero),
                    make_cheb<float>(blas, op, eig_tol, eig_max_it, scale_eig_max, scale_eig_min, max_it, initial_guess_zero));
}
int b200host_cheb3_init_with_ones(void *c) {
    Box *b = (Box *)c;
    if (b->kind != K_CHEB) return 1;
    if (b->real_type == B200FEM_F64)
        std::static_pointer_cast<Chebyshev3<double>>(b->obj)->init_with_ones();
    else
        std::static_pointer_cast<Chebyshev3<float>>(b->obj)->init_with_ones();
    return 0;
}
/* power method from a caller-provided DEVICE guess vector (Chebyshev3::init, sfem_Chebyshev3.hpp:163-174) */
int b200host_cheb3_init(void *c, const void *guess_dev) {
    Box *b = (Box *)c;
    if (b->kind != K_CHEB) return 1;
    if (b->real_type == B200FEM_F64)
        std::static_pointer_cast<Chebyshev3<double>>(b->obj)->init((const double *)guess_dev);
    else
        std::static_pointer_cast<Chebyshev3<float>>(b->obj)->init((const float *)guess_dev);
    return 0;
}
double b200host_cheb3_eig_max(void *c) {
    Box *b = (Box *)c;
    return DISPATCH(b->real_type, (double)std::static_pointer_cast<Chebyshev3<double>>(b->obj)->eig_max,
                    (double)std::static_pointer_cast<Chebyshev3<float>>(b->obj)->eig_max);
}
int b200host_cheb3_power_method_iterations(void *c) {
    Box *b = (Box *)c;
    return DISPATCH(b->real_type, std::static_pointer_cast<Chebyshev3<double>>(b->obj)->power_method_iterations,
                    std::static_pointer_cast<Chebyshev3<float>>(b->obj)->power_method_iterations);
}
int b200host_cheb3_set_initial_guess_zero(void *c, int v) {
    Box *b = (Box *)c;
    if (b->real_type == B200FEM_F64)
        std::static_pointer_cast<Chebyshev3<double>>(b->obj)->set_initial_guess_zero(v != 0);
    else
        std::static_pointer_cast<Chebyshev3<float>>(b->obj)->set_initial_guess_zero(v != 0);
    return 0;
}

void *b200host_mg(void *blas, int max_it, double atol, int cycle_type, int verbose) {
    return DISPATCH(((Box *)blas)->real_type, make_mg<double>(blas, max_it, atol, cycle_type, verbose),
                    make_mg<float>(blas, max_it, atol, cycle_type, verbose));
}
int b200host_mg_add_level(void *mg, void *op, void *smoother, void *prolongation, void *restriction) {
    Box *b = (Box *)mg;
    if (b->kind != K_MG) return 1;
    if (b->real_type == B200FEM_F64)
        std::static_pointer_cast<Multigrid<double>>(b->obj)->add_level(op_of<double>(op), op_of<double>(smoother),
                                                                       op_of<double>(prolongation), op_of<double>(restriction));
    else
        std::static_pointer_cast<Multigrid<float>>(b->obj)->add_level(op_of<float>(op), op_of<float>(smoother),
                                                                      op_of<float>(prolongation), op_of<float>(restriction));
    return 0;
}

/* ---- mesh folders in the reference's on-disk format (b200fem_mesh_io.hpp) ----------------------------------------- */
static thread_local std::string g_mesh_error;
const char *b200host_mesh_last_error() { return g_mesh_error.c_str(); }
void *b200host_mesh_read(const char *folder) {
    auto m = new RawMesh();
    if (read_mesh(folder, *m, &g_mesh_error)) {
        delete m;
        return nullptr;
    }
    return m;
}
void        b200host_mesh_destroy(void *m) { delete (RawMesh *)m; }
ptrdiff_t   b200host_mesh_n_elements(void *m) { return ((RawMesh *)m)->n_elements(); }
ptrdiff_t   b200host_mesh_n_nodes(void *m) { return ((RawMesh *)m)->n_nodes(); }
int         b200host_mesh_nxe(void *m) { return ((RawMesh *)m)->nxe(); }
int         b200host_mesh_dim(void *m) { return (int)((RawMesh *)m)->points.size(); }
const char *b200host_mesh_element_type(void *m) { return ((RawMesh *)m)->element_type.c_str(); }
const int32_t *b200host_mesh_elements(void *m, int corner) { return ((RawMesh *)m)->elements[(size_t)corner].data(); }
const float   *b200host_mesh_points(void *m, int d) { return ((RawMesh *)m)->points[(size_t)d].data(); }
int b200host_mesh_write(const char *folder, const char *element_type, int nxe, ptrdiff_t ne, ptrdiff_t nn, int dim,
                        const int32_t *const *elements, const float *const *points) {
    RawMesh m;
    m.element_type      = element_type;
    m.spatial_dimension = dim;
    for (int k = 0; k < nxe; ++k) m.elements.emplace_back(elements[k], elements[k] + ne);
    for (int d = 0; d < dim; ++d) m.points.emplace_back(points[d], points[d] + nn);
    return write_mesh(folder, m, &g_mesh_error);
}

/* run any object: solvers take (b, x) and update x in place; operators accumulate y += A x */
int b200host_apply(void *obj, const void *b, void *x) {
    Box *o = (Box *)obj;
    return DISPATCH(o->real_type, op_of<double>(o)->apply((const double *)b, (double *)x),
                    op_of<float>(o)->apply((const float *)b, (float *)x));
}
int b200host_apply_overwrite(void *obj, const void *x, void *y) {
    Box *o = (Box *)obj;
    return DISPATCH(o->real_type, op_of<double>(o)->apply_overwrite((const double *)x, (double *)y),
                    op_of<float>(o)->apply_overwrite((const float *)x, (float *)y));
}
ptrdiff_t b200host_rows(void *obj) {
    Box *o = (Box *)obj;
    return DISPATCH(o->real_type, op_of<double>(o)->rows(), op_of<float>(o)->rows());
}
int b200host_iterations(void *obj) {
    Box *o = (Box *)obj;
    switch (o->kind) {
        case K_CG:
            return DISPATCH(o->real_type, std::static_pointer_cast<ConjugateGradient<double>>(o->obj)->iterations(),
                            std::static_pointer_cast<ConjugateGradient<float>>(o->obj)->iterations());
        case K_CHEB:
            return DISPATCH(o->real_type, std::static_pointer_cast<Chebyshev3<double>>(o->obj)->iterations(),
                            std::static_pointer_cast<Chebyshev3<float>>(o->obj)->iterations());
        case K_STATIONARY:
            return DISPATCH(o->real_type, std::static_pointer_cast<StationaryIteration<double>>(o->obj)->iterations(),
                            std::static_pointer_cast<StationaryIteration<float>>(o->obj)->iterations());
        case K_MG:
            return DISPATCH(o->real_type, std::static_pointer_cast<Multigrid<double>>(o->obj)->iterations(),
                            std::static_pointer_cast<Multigrid<float>>(o->obj)->iterations());
        default:
            return -1;
    }
}
/* residual norms recorded by CG (after every iteration, [0] = initial) or MG (fine level, every cycle);
 * returns the number of entries available, copies at most `cap` */
int b200host_residual_history(void *obj, double *out, int cap) {
    Box *o = (Box *)obj;
    if (o->kind == K_CG)
        return DISPATCH(o->real_type,
                        copy_history<double>(std::static_pointer_cast<ConjugateGradient<double>>(o->obj)->residual_history, out, cap),
                        copy_history<float>(std::static_pointer_cast<ConjugateGradient<float>>(o->obj)->residual_history, out, cap));
    if (o->kind == K_MG)
        return DISPATCH(o->real_type, copy_history<double>(std::static_pointer_cast<Multigrid<double>>(o->obj)->residual_history, out, cap),
                        copy_history<float>(std::static_pointer_cast<Multigrid<float>>(o->obj)->residual_history, out, cap));
    return -1;
}
void b200host_destroy(void *obj) { delete (Box *)obj; }
}
