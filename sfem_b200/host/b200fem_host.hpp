/* b200fem_host.hpp -- C++ host layer above the C ABI (include/b200fem.h).
 *
 * Mirrors, for the matrix-free hot path only, the interfaces the reference's drivers program against, so that a
 * user of sfem finds the same names, argument meaning and error behaviour (return 0 = SFEM_SUCCESS, 1 = failure):
 *
 *   Operator<T>, LambdaOperator, make_op          algebra/sfem_Operator.hpp:12-33, 121-152
 *   BLAS<T> (13 virtuals)                          algebra/sfem_tpl_blas.hpp:11-35   -> DeviceBLAS<T> over b200fem_blas_*
 *   Function (ops + Dirichlet constraints)         frontend/sfem_Function.cpp:481-496, 561-600
 *   ConjugateGradient<T>                           algebra/sfem_cg.hpp:80-445 (same recurrences and stopping rules)
 *   PowerMethod<T>                                 algebra/sfem_PowerMethod.hpp:18-55
 *   Chebyshev3<T>                                  algebra/sfem_Chebyshev3.hpp:54-300
 *   StationaryIteration<T>                         algebra/sfem_Stationary.hpp:18-80
 *   ShiftableJacobi<T>, ShiftableBlockSymJacobi<T> algebra/sfem_ShiftableJacobi.hpp:17-144
 *   Multigrid<T>                                   algebra/sfem_Multigrid.hpp:26-346
 *   Communicator, ParallelOperator<T>              algebra/parallel/sfem_ParallelOperator.hpp:13-25 (smesh::Communicator)
 *   ParallelMatrixFreeOperator<T>                  frontend/sfem_ParallelMatrixFreeOperator.cpp:35-105
 *   ParallelConjugateGradient<T>                   algebra/parallel/sfem_parallel_cg.hpp:257-400
 *
 * Everything here is orchestration: every vector and every operator action runs in the CUDA kernels behind the
 * C ABI. There is no host arithmetic on vectors and no CPU fallback. Device-side differences from the reference
 * (all result-preserving):
 *   - operators that can write `y = A x` directly (b200fem_op_apply_overwrite) are used in that form instead of
 *     zeros(y) + apply (the reference's `apply_overwrites_output` switch, sfem_cg.hpp:101-103);
 *   - the CG / Chebyshev vector updates use the fused kernels (the reference fuses them on the host only,
 *     sfem_cg.hpp:23-74, sfem_Chebyshev3.hpp:35-52). */
#ifndef B200FEM_HOST_HPP
#define B200FEM_HOST_HPP

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <vector>

#include "b200fem.h"

namespace b200fem {

    enum ExecutionSpace { EXECUTION_SPACE_HOST = 0, EXECUTION_SPACE_DEVICE = 1, EXECUTION_SPACE_INVALID = -1 };

    template <typename T>
    struct RealType;
    template <>
    struct RealType<double> {
        static constexpr int value = B200FEM_F64;
    };
    template <>
    struct RealType<float> {
        static constexpr int value = B200FEM_F32;
    };

#define B200FEM_TRY(call)                                                                                    \
    do {                                                                                                     \
        if ((call) != B200FEM_SUCCESS) {                                                                     \
            std::fprintf(stderr, "[b200fem] %s failed: %s (%s:%d)\n", #call, b200fem_last_error(), __FILE__, \
                         __LINE__);                                                                          \
            return B200FEM_FAILURE;                                                                          \
        }                                                                                                    \
    } while (0)

    /* ---------------------------------------------------------------------------------------------------- */
    /* BLAS<T>                                                                                              */
    /* ---------------------------------------------------------------------------------------------------- */
    template <typename T>
    class BLAS {
    public:
        virtual ~BLAS()                                                                                   = default;
        virtual T   *allocate(const std::size_t n)                                                        = 0;
        virtual void destroy(void *a)                                                                     = 0;
        virtual void zeros(const std::size_t n, T *const x)                                               = 0;
        virtual void values(const std::size_t n, const T value, T *const x)                               = 0;
        virtual void copy(const ptrdiff_t n, const T *const src, T *const dest)                           = 0;
        virtual T    dot(const ptrdiff_t n, const T *const l, const T *const r)                           = 0;
        virtual void axpy(const ptrdiff_t n, const T alpha, const T *const x, T *const y)                 = 0;
        virtual void axpby(const ptrdiff_t n, const T alpha, const T *const x, const T beta, T *const y)  = 0;
        virtual void scal(const ptrdiff_t n, const T alpha, T *const x)                                   = 0;
        virtual void reciprocal(const ptrdiff_t n, const T alpha, T *const x)                             = 0;
        virtual T    norm2(const ptrdiff_t n, const T *const x)                                           = 0;
        virtual void zaxpby(const ptrdiff_t n, const T alpha, const T *const x, const T beta, const T *const y,
                            T *const z)                                                                   = 0;
        virtual void xypaz(const ptrdiff_t n, const T *const x, const T *const y, const T alpha, T *const z) = 0;
        virtual bool good() const { return true; }

        /* fused solver updates (cg_host_:: / chebyshev3_host_:: of the reference) */
        virtual T    cg_update_x_r_and_rtr(const ptrdiff_t n, const T alpha, const T *p, const T *Ap, T *x, T *r) = 0;
        virtual void cg_update_p(const ptrdiff_t n, const T beta, const T *z, T *p)                               = 0;
        virtual void cheb3_iteration_update(const ptrdiff_t n, const T alpha, const T beta, const T *rhs, const T *Ax,
                                            T *p, T *x)                                                           = 0;
    };

    /* Scalars are reduced over ranks through `allreduce_sum` when set (ParallelConjugateGradient,
     * algebra/parallel/sfem_parallel_cg.hpp:169-172: local dot over owned entries, then comm->sum). */
    template <typename T>
    class DeviceBLAS final : public BLAS<T> {
    public:
        static constexpr int            RT = RealType<T>::value;
        void                           *stream{nullptr};
        std::function<double(double)>   allreduce_sum;
        long                            n_launches{0};
        long                            n_reductions{0};

        /* Distributed reductions without leaving C++: every dot / norm / fused CG reduction stays on the device, is
         * summed over the ranks there (ncclAllReduce on one double) and read back once. Takes precedence over
         * `allreduce_sum`. */
        b200fem_comm_t comm{nullptr};
        ~DeviceBLAS() override {
            if (scalar_dev_) b200fem_free(scalar_dev_);
        }
        bool distributed() const { return comm && b200fem_comm_size(comm) > 1; }

        /* ParallelOperator::row_allocation_size(): a vector of `n_owned` entries is allocated with `n_local` entries
         * so that the halo gather can fill its ghost block in place (multigrid levels register one pair each). */
        void reserve_allocation(std::size_t n_owned, std::size_t n_local) {
            auto it = alloc_size_.find(n_owned);
            if (it == alloc_size_.end() || it->second < n_local) alloc_size_[n_owned] = n_local;
        }
        std::size_t allocation_size(std::size_t n) const {
            auto it = alloc_size_.find(n);
            return (it == alloc_size_.end() || it->second < n) ? n : it->second;
        }

        /* Optional allocator: distributed solvers need work vectors with room for the ghost block
         * (ParallelOperator::row_allocation_size(), algebra/parallel/sfem_ParallelOperator.hpp:13-25) and, for the
         * owner-push halo exchange, vectors that live in NVLink-mapped symmetric memory. Must return zero-filled
         * device memory of at least `bytes`. */
        std::function<void *(std::size_t bytes)> alloc_hook;
        std::function<void(void *)>              free_hook;

        T *allocate(const std::size_t n_requested) override {
            void             *p = nullptr;
            const std::size_t n = allocation_size(n_requested);
            if (alloc_hook) {
                p = alloc_hook((n ? n : 1) * sizeof(T));
                if (!p) ok_ = false;
                return (T *)p;
            }
            if (b200fem_malloc(&p, (n ? n : 1) * sizeof(T))) {
                std::fprintf(stderr, "[b200fem] allocate(%zu) failed: %s\n", n, b200fem_last_error());
                ok_ = false;
                return nullptr;
            }
            return (T *)p;
        }
        void destroy(void *a) override {
            if (free_hook)
                free_hook(a);
            else
                b200fem_free(a);
        }
        void zeros(const std::size_t n, T *const x) override { track(b200fem_blas_zeros(RT, (ptrdiff_t)n, x, stream)); }
        void values(const std::size_t n, const T v, T *const x) override {
            track(b200fem_blas_values(RT, (ptrdiff_t)n, (double)v, x, stream));
        }
        void copy(const ptrdiff_t n, const T *const s, T *const d) override { track(b200fem_blas_copy(RT, n, s, d, stream)); }
        T    dot(const ptrdiff_t n, const T *const l, const T *const r) override {
            double v = 0;
            if (distributed()) {
                track(b200fem_blas_dot_async(RT, n, l, r, scalar_dev(), stream));
                return (T)sum_over_ranks();
            }
            track(b200fem_blas_dot(RT, n, l, r, &v, stream));
            return (T)reduce(v);
        }
        void axpy(const ptrdiff_t n, const T a, const T *const x, T *const y) override {
            track(b200fem_blas_axpy(RT, n, (double)a, x, y, stream));
        }
        void axpby(const ptrdiff_t n, const T a, const T *const x, const T b, T *const y) override {
            track(b200fem_blas_axpby(RT, n, (double)a, x, (double)b, y, stream));
        }
        void scal(const ptrdiff_t n, const T a, T *const x) override { track(b200fem_blas_scal(RT, n, (double)a, x, stream)); }
        void reciprocal(const ptrdiff_t n, const T a, T *const x) override {
            track(b200fem_blas_reciprocal(RT, n, (double)a, x, stream));
        }
        T norm2(const ptrdiff_t n, const T *const x) override {
            if (distributed()) return (T)std::sqrt((double)dot(n, x, x));
            if (!allreduce_sum) {
                double v = 0;
                track(b200fem_blas_norm2(RT, n, x, &v, stream));
                return (T)v;
            }
            double v = 0;
            track(b200fem_blas_dot(RT, n, x, x, &v, stream));
            return (T)std::sqrt(reduce(v));
        }
        void zaxpby(const ptrdiff_t n, const T a, const T *const x, const T b, const T *const y, T *const z) override {
            track(b200fem_blas_zaxpby(RT, n, (double)a, x, (double)b, y, z, stream));
        }
        void xypaz(const ptrdiff_t n, const T *const x, const T *const y, const T a, T *const z) override {
            track(b200fem_blas_xypaz(RT, n, x, y, (double)a, z, stream));
        }
        bool good() const override { return ok_; }

        T cg_update_x_r_and_rtr(const ptrdiff_t n, const T alpha, const T *p, const T *Ap, T *x, T *r) override {
            double v = 0;
            if (distributed()) {
                track(b200fem_cg_update_x_r_rtr_async(RT, n, (double)alpha, p, Ap, x, r, scalar_dev(), stream));
                return (T)sum_over_ranks();
            }
            track(b200fem_cg_update_x_r_rtr(RT, n, (double)alpha, p, Ap, x, r, &v, stream));
            return (T)reduce(v);
        }
        void cg_update_p(const ptrdiff_t n, const T beta, const T *z, T *p) override {
            track(b200fem_cg_update_p(RT, n, (double)beta, z, p, stream));
        }
        void cheb3_iteration_update(const ptrdiff_t n, const T alpha, const T beta, const T *rhs, const T *Ax, T *p,
                                    T *x) override {
            track(b200fem_cheb3_update(RT, n, (double)alpha, (double)beta, rhs, Ax, p, x, stream));
        }

    private:
        bool                               ok_{true};
        double                            *scalar_dev_{nullptr};
        std::map<std::size_t, std::size_t> alloc_size_;
        double                            *scalar_dev() {
            if (!scalar_dev_) {
                void *p = nullptr;
                if (b200fem_malloc(&p, 64 * sizeof(double))) ok_ = false;
                scalar_dev_ = (double *)p;
            }
            return scalar_dev_;
        }
        double sum_over_ranks() {
            double v = 0;
            n_reductions++;
            if (b200fem_comm_sum_to_host(comm, scalar_dev_, 1, &v, stream)) {
                std::fprintf(stderr, "[b200fem] distributed reduction failed: %s\n", b200fem_last_error());
                ok_ = false;
            }
            return v;
        }
        double reduce(double v) const { return allreduce_sum ? allreduce_sum(v) : v; }
        void   track(int rc) {
            n_launches++;
            if (rc != B200FEM_SUCCESS) {
                std::fprintf(stderr, "[b200fem] vector kernel failed: %s\n", b200fem_last_error());
                ok_ = false;
            }
        }
    };

    template <typename T>
    std::shared_ptr<DeviceBLAS<T>> make_device_blas() {
        return std::make_shared<DeviceBLAS<T>>();
    }

    /* RAII device vector allocated through a BLAS (zero-filled, like BLAS<T>::allocate) */
    template <typename T>
    class Buffer {
    public:
        Buffer(std::size_t n, const std::shared_ptr<BLAS<T>> &blas) : n_(n), blas_(blas), ptr_(blas->allocate(n)) {}
        Buffer(std::size_t n, T *wrapped) : n_(n), ptr_(wrapped) {}
        ~Buffer() {
            if (blas_ && ptr_) blas_->destroy(ptr_);
        }
        Buffer(const Buffer &)            = delete;
        Buffer &operator=(const Buffer &) = delete;
        T          *data() const { return ptr_; }
        std::size_t size() const { return n_; }

    private:
        std::size_t              n_;
        std::shared_ptr<BLAS<T>> blas_;
        T                       *ptr_;
    };
    template <typename T>
    using SharedBuffer = std::shared_ptr<Buffer<T>>;
    template <typename T>
    SharedBuffer<T> create_buffer(std::size_t n, const std::shared_ptr<BLAS<T>> &blas) {
        return std::make_shared<Buffer<T>>(n, blas);
    }

    /* ---------------------------------------------------------------------------------------------------- */
    /* Operator<T>                                                                                          */
    /* ---------------------------------------------------------------------------------------------------- */
    template <typename T>
    class Operator {
    public:
        virtual ~Operator()                             = default;
        /* y += A x for element operators; solvers and smoothers update x in place from (b, x) */
        virtual int            apply(const T *const x, T *const y) = 0;
        virtual std::ptrdiff_t rows() const                        = 0;
        virtual std::ptrdiff_t cols() const                        = 0;
        virtual ExecutionSpace execution_space() const { return EXECUTION_SPACE_DEVICE; }
        /* y = A x without a preceding zeros(y); only when can_overwrite() */
        virtual bool can_overwrite() const { return false; }
        virtual int  apply_overwrite(const T *const, T *const) { return B200FEM_FAILURE; }
        /* The overwriting apply over the element sub-range [e0, e1) (b200fem_op_apply_range): `first` zeroes what the whole
         * apply zeroes, `last` finishes the apply (constrained dofs). range_elements() = the number of elements when the
         * operator can be applied in pieces, 0 otherwise. Used by the distributed apply to run the ghost-independent
         * elements between the two halves of the halo gather. */
        virtual std::ptrdiff_t range_elements() const { return 0; }
        /* y = A x in one launch whose elements >= e_dependent wait inside the kernel for *flag >= want
         * (b200fem_op_apply_overwrite_after); can_wait_in_kernel() tells whether the operator has that form */
        virtual bool can_wait_in_kernel() const { return false; }
        virtual int  apply_overwrite_after(const T *const, T *const, std::ptrdiff_t, const unsigned long long *, unsigned long long) {
            return B200FEM_FAILURE;
        }
        virtual int apply_overwrite_range(const T *const, T *const, std::ptrdiff_t, std::ptrdiff_t, bool /*first*/, bool /*last*/) {
            return B200FEM_FAILURE;
        }
    };

    template <typename T>
    class LambdaOperator final : public Operator<T> {
    public:
        using Fn      = std::function<int(const T *const, T *const)>;
        using RangeFn = std::function<int(const T *const, T *const, std::ptrdiff_t, std::ptrdiff_t, bool, bool)>;
        LambdaOperator(std::ptrdiff_t rows, std::ptrdiff_t cols, Fn apply, Fn overwrite = nullptr)
            : rows_(rows), cols_(cols), apply_(std::move(apply)), overwrite_(std::move(overwrite)) {}
        void set_range(std::function<std::ptrdiff_t()> n_elements, RangeFn fn) { /* the count is asked for at apply time */
            range_elements_ = std::move(n_elements);
            range_          = std::move(fn);
        }
        std::ptrdiff_t range_elements() const override { return (range_ && range_elements_) ? range_elements_() : 0; }
        using AfterFn = std::function<int(const T *const, T *const, std::ptrdiff_t, const unsigned long long *, unsigned long long)>;
        void set_after(std::function<bool()> can, AfterFn fn) {
            can_after_ = std::move(can);
            after_     = std::move(fn);
        }
        bool can_wait_in_kernel() const override { return after_ && can_after_ && can_after_(); }
        int  apply_overwrite_after(const T *const x, T *const y, std::ptrdiff_t e_dep, const unsigned long long *flag,
                                   unsigned long long want) override {
            return after_ ? after_(x, y, e_dep, flag, want) : B200FEM_FAILURE;
        }
        int apply_overwrite_range(const T *const x, T *const y, std::ptrdiff_t e0, std::ptrdiff_t e1, bool first, bool last) override {
            return range_ ? range_(x, y, e0, e1, first, last) : B200FEM_FAILURE;
        }
        int            apply(const T *const x, T *const y) override { return apply_(x, y); }
        std::ptrdiff_t rows() const override { return rows_; }
        std::ptrdiff_t cols() const override { return cols_; }
        bool           can_overwrite() const override { return (bool)overwrite_; }
        int            apply_overwrite(const T *const x, T *const y) override { return overwrite_(x, y); }

    private:
        std::ptrdiff_t rows_, cols_;
        Fn             apply_, overwrite_;
        RangeFn        range_;
        std::function<std::ptrdiff_t()> range_elements_;
        std::function<bool()>           can_after_;
        AfterFn                         after_;
    };

    template <typename T>
    std::shared_ptr<Operator<T>> make_op(std::ptrdiff_t rows, std::ptrdiff_t cols, typename LambdaOperator<T>::Fn fn,
                                         typename LambdaOperator<T>::Fn overwrite = nullptr) {
        return std::make_shared<LambdaOperator<T>>(rows, cols, std::move(fn), std::move(overwrite));
    }

    /* zeros(y) + apply, or the overwriting form when the operator has one */
    template <typename T>
    int apply_into(Operator<T> &op, BLAS<T> &blas, const T *x, T *y) {
        if (op.can_overwrite()) return op.apply_overwrite(x, y);
        blas.zeros(op.rows(), y);
        return op.apply(x, y);
    }

    /* ---------------------------------------------------------------------------------------------------- */
    /* Function: operators + Dirichlet conditions on one space (frontend/sfem_Function.cpp)                 */
    /* ---------------------------------------------------------------------------------------------------- */
    struct DirichletCondition {
        const int32_t *idx_dev{nullptr}; /* device list of node ids */
        ptrdiff_t      n{0};
        int            component{0};
        double         value{0};
    };

    template <typename T>
    class Function {
    public:
        static constexpr int RT = RealType<T>::value;
        void                *stream{nullptr};

        Function(ptrdiff_t n_nodes, int block_size) : n_nodes_(n_nodes), block_size_(block_size) {}

        void add_operator(b200fem_op_t op) { ops_.push_back(op); } /* not owned */
        void add_constraint(const DirichletCondition &c) { constraints_.push_back(c); }
        ptrdiff_t n_dofs() const { return n_nodes_ * block_size_; }
        int       block_size() const { return block_size_; }
        const std::vector<DirichletCondition> &constraints() const { return constraints_; }

        /* out += sum_op A_op h ; out[c] = h[c] on constrained dofs   (Function::apply :481-496) */
        int apply(const T * /*x*/, const T *const h, T *const out) {
            for (auto op : ops_) B200FEM_TRY(b200fem_op_apply(op, h, out, stream));
            return copy_constrained_dofs(h, out);
        }
        int apply_overwrite(const T *const h, T *const out) {
            if (ops_.empty()) return B200FEM_FAILURE;
            B200FEM_TRY(b200fem_op_apply_overwrite(ops_[0], h, out, stream));
            for (size_t i = 1; i < ops_.size(); i++) B200FEM_TRY(b200fem_op_apply(ops_[i], h, out, stream));
            return copy_constrained_dofs(h, out);
        }
        /* the overwriting apply in pieces (one operator only): element range [e0, e1) */
        std::ptrdiff_t range_elements() const { return ops_.size() == 1 ? b200fem_op_nelements(ops_[0]) : 0; }
        int apply_overwrite_range(const T *const h, T *const out, std::ptrdiff_t e0, std::ptrdiff_t e1, bool first, bool last) {
            if (ops_.size() != 1) return B200FEM_FAILURE;
            B200FEM_TRY(b200fem_op_apply_range(ops_[0], h, out, e0, e1, first ? 1 : 2, stream));
            return last ? copy_constrained_dofs(h, out) : B200FEM_SUCCESS;
        }
        bool can_wait_in_kernel() const {
            int v = 0;
            return ops_.size() == 1 && b200fem_op_get_option(ops_[0], "ghost_wait", &v) == B200FEM_SUCCESS && v == 1;
        }
        int apply_overwrite_after(const T *const h, T *const out, std::ptrdiff_t e_dep, const unsigned long long *flag, unsigned long long want) {
            if (ops_.size() != 1) return B200FEM_FAILURE;
            B200FEM_TRY(b200fem_op_apply_overwrite_after(ops_[0], h, out, e_dep, flag, want, stream));
            return copy_constrained_dofs(h, out);
        }
        int hessian_diag(const T * /*x*/, T *const values) {
            for (auto op : ops_) B200FEM_TRY(b200fem_op_diag(op, values, stream));
            return B200FEM_SUCCESS;
        }
        int hessian_block_diag_sym(const T * /*x*/, T *const values) {
            for (auto op : ops_) B200FEM_TRY(b200fem_op_block_diag_sym(op, values, stream));
            return B200FEM_SUCCESS;
        }
        int apply_constraints(T *const x) { /* x[c] = value (:561-577) */
            for (auto &c : constraints_)
                B200FEM_TRY(b200fem_constraint_nodes_to_value_vec(RT, c.n, c.idx_dev, block_size_, c.component, c.value, x,
                                                                  stream));
            return B200FEM_SUCCESS;
        }
        int set_value_to_constrained_dofs(const T val, T *const x) { /* (:588-596) */
            for (auto &c : constraints_)
                B200FEM_TRY(b200fem_constraint_nodes_to_value_vec(RT, c.n, c.idx_dev, block_size_, c.component, (double)val, x,
                                                                  stream));
            return B200FEM_SUCCESS;
        }
        int apply_zero_constraints(T *const x) { return set_value_to_constrained_dofs(T(0), x); } /* (:579-586) */
        int copy_constrained_dofs(const T *const src, T *const dest) {
            for (auto &c : constraints_)
                B200FEM_TRY(b200fem_constraint_nodes_copy_vec(RT, c.n, c.idx_dev, block_size_, c.component, src, dest,
                                                              stream));
            return B200FEM_SUCCESS;
        }
        int launches_per_apply(bool overwrite) const {
            int n = (int)constraints_.size();
            for (size_t i = 0; i < ops_.size(); i++) n += b200fem_op_gpu_launches_per_apply(ops_[i], overwrite && i == 0);
            return n;
        }

        /* make_linear_op(f) (frontend/sfem_API.hpp): Operator<T> over Function::apply(nullptr, x, y) */
        std::shared_ptr<Operator<T>> linear_op(const std::shared_ptr<Function<T>> &self) {
            auto op = std::make_shared<LambdaOperator<T>>(
                    n_dofs(), n_dofs(), [self](const T *const x, T *const y) { return self->apply(nullptr, x, y); },
                    [self](const T *const x, T *const y) { return self->apply_overwrite(x, y); });
            op->set_after([self]() { return self->can_wait_in_kernel(); },
                          [self](const T *const x, T *const y, std::ptrdiff_t e_dep, const unsigned long long *flag, unsigned long long want) {
                              return self->apply_overwrite_after(x, y, e_dep, flag, want);
                          });
            op->set_range([self]() { return self->range_elements(); }, [self](const T *const x, T *const y, std::ptrdiff_t e0, std::ptrdiff_t e1, bool first, bool last) {
                return self->apply_overwrite_range(x, y, e0, e1, first, last);
            });
            return op;
        }

    private:
        ptrdiff_t                       n_nodes_;
        int                             block_size_;
        std::vector<b200fem_op_t>       ops_;
        std::vector<DirichletCondition> constraints_;
    };

    /* ---------------------------------------------------------------------------------------------------- */
    /* ConjugateGradient                                                                                    */
    /* ---------------------------------------------------------------------------------------------------- */
    template <typename T>
    class MatrixFreeLinearSolver : public Operator<T> {
    public:
        virtual void set_op(const std::shared_ptr<Operator<T>> &op)                = 0;
        virtual void set_preconditioner_op(const std::shared_ptr<Operator<T>> &op) = 0;
        virtual void set_max_it(const int it)                                      = 0;
        virtual void set_n_dofs(const ptrdiff_t n)                                 = 0;
        virtual void set_initial_guess_zero(const bool) {}
        virtual int  iterations() const = 0;
    };

    template <typename T>
    class ConjugateGradient final : public MatrixFreeLinearSolver<T> {
    public:
        std::shared_ptr<Operator<T>> apply_op;
        std::shared_ptr<Operator<T>> preconditioner_op;
        std::shared_ptr<BLAS<T>>     blas;

        T         rtol{1e-10};
        T         atol{1e-16};
        int       max_it{10000};
        int       check_each{1};
        ptrdiff_t n_dofs{-1};
        int       iterations_{0};
        bool      verbose{false};
        /* residual norm after every iteration ([0] = initial); what the reference prints in monitor() */
        std::vector<double> residual_history;

        int            iterations() const override { return iterations_; }
        void           set_atol(const T v) { atol = v; }
        void           set_rtol(const T v) { rtol = v; }
        void           set_verbose(const bool v) { verbose = v; }
        std::ptrdiff_t rows() const override { return n_dofs; }
        std::ptrdiff_t cols() const override { return n_dofs; }
        void           set_op(const std::shared_ptr<Operator<T>> &op) override {
            apply_op = op;
            n_dofs   = op->rows();
        }
        void set_preconditioner_op(const std::shared_ptr<Operator<T>> &op) override { preconditioner_op = op; }
        void set_max_it(const int it) override { max_it = it; }
        void set_n_dofs(const ptrdiff_t n) override { n_dofs = n; }
        bool good() const { return blas && blas->good() && apply_op; }

        int apply(const T *const b, T *const x) override {
            if (n_dofs < 0) {
                std::fprintf(stderr, "Error uninitiaized n_dofs. Set set_n_dofs to set the number of dofs\n");
                return B200FEM_FAILURE;
            }
            if (!good()) return B200FEM_FAILURE;
            ensure_workspace(n_dofs, (bool)preconditioner_op);
            residual_history.clear();
            return preconditioner_op ? solve_preconditioned(n_dofs, b, x) : solve_basic(n_dofs, b, x);
        }

    private:
        SharedBuffer<T> r_, z_, p_, Ap_;

        void ensure_workspace(const ptrdiff_t n, const bool need_z) {
            if (!r_ || (ptrdiff_t)r_->size() != n) {
                r_  = create_buffer<T>(n, blas);
                p_  = create_buffer<T>(n, blas);
                Ap_ = create_buffer<T>(n, blas);
                z_.reset();
            }
            if (need_z && !z_) z_ = create_buffer<T>(n, blas);
        }

        void monitor(const int iter, const T residual, const T relative, const T alpha) {
            residual_history.push_back((double)residual);
            if (!verbose) return;
            if (iter == max_it || iter == 0 || iter % check_each == 0 || relative < rtol)
                std::printf("%d: residual abs: %g, rel: %g (rtol = %g, atol = %g, alpha = %g)\n", iter, (double)residual,
                            (double)relative, (double)rtol, (double)atol, (double)alpha);
        }

        /* sfem_cg.hpp:301-363 */
        int solve_basic(const ptrdiff_t n, const T *const b, T *const x) {
            T *r = r_->data(), *p = p_->data(), *Ap = Ap_->data();
            if (apply_into(*apply_op, *blas, x, r)) return B200FEM_FAILURE;
            blas->axpby(n, T(1), b, T(-1), r);

            const T rtr0    = blas->dot(n, r, r);
            const T r_norm0 = std::sqrt(rtr0);
            monitor(0, r_norm0, 1, 0);
            T rtr = rtr0;
            if (rtr0 == 0) return B200FEM_SUCCESS;
            blas->copy(n, r, p);

            int info = B200FEM_FAILURE;
            for (iterations_ = 0; iterations_ < max_it; iterations_++) {
                if (apply_into(*apply_op, *blas, p, Ap)) return B200FEM_FAILURE;
                const T ptAp    = blas->dot(n, p, Ap);
                const T alpha   = rtr / ptAp;
                const T rtr_new = blas->cg_update_x_r_and_rtr(n, alpha, p, Ap, x, r);
                const T beta    = rtr_new / rtr;
                rtr             = rtr_new;
                blas->cg_update_p(n, beta, r, p);

                const T r_norm = std::sqrt(rtr_new);
                monitor(iterations_ + 1, r_norm, r_norm / r_norm0, alpha);
                if (r_norm < atol || rtr_new == 0 || r_norm / r_norm0 < rtol) {
                    info = B200FEM_SUCCESS;
                    break;
                }
            }
            return blas->good() ? info : B200FEM_FAILURE;
        }

        /* sfem_cg.hpp:365-445 */
        int solve_preconditioned(const ptrdiff_t n, const T *const b, T *const x) {
            T *r = r_->data(), *z = z_->data(), *p = p_->data(), *Ap = Ap_->data();
            if (apply_into(*apply_op, *blas, x, r)) return B200FEM_FAILURE;
            blas->axpby(n, T(1), b, T(-1), r);

            const T rtr0    = blas->dot(n, r, r);
            const T r_norm0 = std::sqrt(rtr0);
            monitor(0, r_norm0, 1, 0);
            if (rtr0 == 0) return B200FEM_SUCCESS;

            if (apply_into(*preconditioner_op, *blas, r, z)) return B200FEM_FAILURE;
            blas->copy(n, z, p);
            if (apply_into(*apply_op, *blas, p, Ap)) return B200FEM_FAILURE;

            const T rtz0 = blas->dot(n, r, z);
            T       rtz  = rtz0;
            if (rtz == 0) return B200FEM_SUCCESS;
            {
                const T ptAp  = blas->dot(n, p, Ap);
                const T alpha = rtz / ptAp;
                blas->axpy(n, alpha, p, x);
                blas->axpy(n, -alpha, Ap, r);
            }

            int info = B200FEM_FAILURE;
            for (iterations_ = 0; iterations_ < max_it; iterations_++) {
                if (apply_into(*preconditioner_op, *blas, r, z)) return B200FEM_FAILURE;
                const T rtz_new = blas->dot(n, r, z);
                const T beta    = rtz_new / rtz;
                rtz             = rtz_new;
                blas->cg_update_p(n, beta, z, p);

                if (apply_into(*apply_op, *blas, p, Ap)) return B200FEM_FAILURE;
                const T ptAp  = blas->dot(n, p, Ap);
                const T alpha = rtz / ptAp;
                blas->axpy(n, alpha, p, x);
                blas->axpy(n, -alpha, Ap, r);

                const T anorm = std::sqrt(rtz);
                const T rnorm = anorm / std::sqrt(rtz0);
                monitor(iterations_ + 1, anorm, rnorm, alpha);
                if (anorm < atol || rnorm < rtol) {
                    info = B200FEM_SUCCESS;
                    break;
                }
            }
            return blas->good() ? info : B200FEM_FAILURE;
        }
    };

    /* ---------------------------------------------------------------------------------------------------- */
    /* Jacobi smoothers                                                                                     */
    /* ---------------------------------------------------------------------------------------------------- */
    /* x += (relaxation / d) * b   (ShiftableJacobi::apply, sfem_ShiftableJacobi.hpp:47-52) */
    template <typename T>
    class ShiftableJacobi final : public Operator<T> {
    public:
        std::shared_ptr<BLAS<T>> blas;
        SharedBuffer<T>          inv_diag;
        const T                 *diag{nullptr};
        ptrdiff_t                n{0};
        T                        relaxation_parameter{0.3};

        void set_diag(const T *d, ptrdiff_t n_) { /* d stays caller-owned */
            diag     = d;
            n        = n_;
            inv_diag = create_buffer<T>(n, blas);
            blas->copy(n, diag, inv_diag->data());
            blas->reciprocal(n, relaxation_parameter, inv_diag->data());
        }
        int shift(const T *d) {
            blas->copy(n, diag, inv_diag->data());
            blas->axpy(n, 1, d, inv_diag->data());
            blas->reciprocal(n, relaxation_parameter, inv_diag->data());
            return B200FEM_SUCCESS;
        }
        int apply(const T *const b, T *const x) override {
            blas->xypaz(n, inv_diag->data(), b, 1, x);
            return blas->good() ? B200FEM_SUCCESS : B200FEM_FAILURE;
        }
        std::ptrdiff_t rows() const override { return n; }
        std::ptrdiff_t cols() const override { return n; }
    };

    /* y += relaxation * inv(B_node) x per 3x3 node block (ShiftableBlockSymJacobi, :67-144) */
    template <typename T>
    class ShiftableBlockSymJacobi final : public Operator<T> {
    public:
        static constexpr int     RT = RealType<T>::value;
        std::shared_ptr<BLAS<T>> blas;
        void                    *stream{nullptr};
        SharedBuffer<T>          inv_diag;
        const T                 *diag{nullptr};
        const void              *constraints_mask{nullptr}; /* device, mask_t layout */
        ptrdiff_t                n_blocks{0};
        T                        relaxation_parameter{1. / 3};
        int                      block_size{3};

        int set_diag(const T *sym_diag6, ptrdiff_t n_blocks_) {
            diag     = sym_diag6;
            n_blocks = n_blocks_;
            inv_diag = create_buffer<T>(n_blocks * 9, blas);
            return b200fem_block_sym_jacobi_setup(RT, n_blocks, diag, nullptr, constraints_mask, (double)relaxation_parameter,
                                                  inv_diag->data(), stream);
        }
        int shift(const T *shift6) {
            return b200fem_block_sym_jacobi_setup(RT, n_blocks, diag, shift6, constraints_mask, (double)relaxation_parameter,
                                                  inv_diag->data(), stream);
        }
        int apply(const T *const x, T *const y) override {
            return b200fem_block_jacobi_apply(RT, n_blocks, inv_diag->data(), x, y, stream);
        }
        std::ptrdiff_t rows() const override { return n_blocks * 3; }
        std::ptrdiff_t cols() const override { return n_blocks * 3; }
    };

    /* MixedPrecisionShiftableBlockSymJacobi<double, float> (algebra/sfem_MixedPrecisionShiftableBlockSymJacobi.hpp): fp64
     * vectors and block diagonal, the inverted blocks kept in fp32 */
    class MixedPrecisionShiftableBlockSymJacobi final : public Operator<double> {
    public:
        void         *stream{nullptr};
        const double *diag{nullptr};
        const void   *constraints_mask{nullptr};
        ptrdiff_t     n_blocks{0};
        double        relaxation_parameter{1. / 3};
        ~MixedPrecisionShiftableBlockSymJacobi() override {
            if (inv_diag_) b200fem_free(inv_diag_);
        }
        int set_diag(const double *sym_diag6, ptrdiff_t n_blocks_) {
            diag     = sym_diag6;
            n_blocks = n_blocks_;
            if (inv_diag_) b200fem_free(inv_diag_);
            void *p = nullptr;
            B200FEM_TRY(b200fem_malloc(&p, (size_t)(n_blocks ? n_blocks : 1) * 9 * sizeof(float)));
            inv_diag_ = (float *)p;
            return b200fem_block_sym_jacobi_setup_mixed(n_blocks, diag, nullptr, constraints_mask, relaxation_parameter, inv_diag_, stream);
        }
        int shift(const double *shift6) {
            return b200fem_block_sym_jacobi_setup_mixed(n_blocks, diag, shift6, constraints_mask, relaxation_parameter, inv_diag_, stream);
        }
        int apply(const double *const x, double *const y) override {
            return b200fem_block_jacobi_apply_mixed(n_blocks, inv_diag_, x, y, stream);
        }
        std::ptrdiff_t rows() const override { return n_blocks * 3; }
        std::ptrdiff_t cols() const override { return n_blocks * 3; }

    private:
        float *inv_diag_{nullptr};
    };

    /* x <- x + P (b - A x), max_it times (StationaryIteration::apply, sfem_Stationary.hpp:50-80) */
    template <typename T>
    class StationaryIteration final : public MatrixFreeLinearSolver<T> {
    public:
        ptrdiff_t                    n_dofs{-1};
        int                          max_it{3};
        std::shared_ptr<Operator<T>> op, preconditioner;
        std::shared_ptr<BLAS<T>>     blas;
        bool                         use_arg_as_first_residual{false};
        int                          iterations_{0};

        int  iterations() const override { return iterations_; }
        void set_op(const std::shared_ptr<Operator<T>> &o) override {
            op     = o;
            n_dofs = o->rows();
        }
        void set_preconditioner_op(const std::shared_ptr<Operator<T>> &o) override { preconditioner = o; }
        void set_max_it(const int it) override { max_it = it; }
        void set_n_dofs(const ptrdiff_t n) override { n_dofs = n; }
        std::ptrdiff_t rows() const override { return n_dofs; }
        std::ptrdiff_t cols() const override { return n_dofs; }

        int apply(const T *const b, T *const x) override {
            iterations_ = 0;
            if (use_arg_as_first_residual) {
                iterations_ = 1;
                if (preconditioner->apply(b, x)) return B200FEM_FAILURE;
            }
            if (iterations_ >= max_it) return B200FEM_SUCCESS;
            if (!work_ || (ptrdiff_t)work_->size() != n_dofs) work_ = create_buffer<T>(n_dofs, blas);
            T *r = work_->data();
            for (; iterations_ < max_it; iterations_++) {
                if (apply_into(*op, *blas, x, r)) return B200FEM_FAILURE;
                blas->axpby(n_dofs, T(1), b, T(-1), r);
                if (preconditioner->apply(r, x)) return B200FEM_FAILURE;
            }
            return blas->good() ? B200FEM_SUCCESS : B200FEM_FAILURE;
        }

    private:
        SharedBuffer<T> work_;
    };

    /* ---------------------------------------------------------------------------------------------------- */
    /* PowerMethod + Chebyshev3                                                                             */
    /* ---------------------------------------------------------------------------------------------------- */
    template <typename T>
    class PowerMethod {
    public:
        std::shared_ptr<BLAS<T>> blas;
        int                      iterations{0};
        bool                     verbose{false};

        /* same sequence of operations as sfem_PowerMethod.hpp:18-55 */
        T max_eigen_value(Operator<T> &op, const int max_it, const T stol, const ptrdiff_t n, T *const eigen_vector,
                          T *const work) {
            const int max_it_odd = max_it + (max_it % 2 == 0);
            apply_into(op, *blas, eigen_vector, work);
            blas->scal(n, 1 / blas->norm2(n, work), work);
            T *vecs[2]    = {eigen_vector, work};
            T  lambda     = 0;
            T  lambda_old = 0;
            for (int k = 0; k < max_it_odd; k++) {
                T *w  = vecs[k % 2];
                T *ev = vecs[(k + 1) % 2];
                apply_into(op, *blas, ev, w);
                lambda = blas->norm2(n, w);
                blas->scal(n, 1 / lambda, w);
                iterations = k + 1;
                if (std::abs(lambda - lambda_old) <= stol && k % 2 == 0) {
                    if (verbose) std::printf("PowerMethod (%d): lambda = %g\n", k, (double)lambda);
                    break;
                }
                lambda_old = lambda;
            }
            return lambda;
        }
    };

    template <typename T>
    class Chebyshev3 final : public MatrixFreeLinearSolver<T> {
    public:
        std::shared_ptr<Operator<T>> apply_op;
        std::shared_ptr<BLAS<T>>     blas;
        SharedBuffer<T>              p_, temp_;

        T         atol{1e-10};
        T         rtol{1e-10};
        T         eigen_solver_tol{1e-6};
        int       eigen_solver_max_it{1000};
        int       max_it{3};
        int       iterations_{0};
        T         eig_max{0};
        T         scale_eig_max{1};
        T         scale_eig_min{0.06};
        ptrdiff_t n_dofs{-1};
        bool      is_initial_guess_zero{false};
        int       power_method_iterations{0};

        int  iterations() const override { return iterations_; }
        void set_initial_guess_zero(const bool v) override { is_initial_guess_zero = v; }
        void set_op(const std::shared_ptr<Operator<T>> &op) override {
            n_dofs   = op->rows();
            apply_op = op;
        }
        void set_preconditioner_op(const std::shared_ptr<Operator<T>> &) override {}
        void set_n_dofs(const ptrdiff_t n) override { n_dofs = n; }
        void set_max_it(const int it) override { max_it = it; }
        std::ptrdiff_t rows() const override { return n_dofs; }
        std::ptrdiff_t cols() const override { return n_dofs; }
        bool           good() const { return blas && blas->good() && apply_op; }

        void init_with_ones() { /* :138-146 */
            auto ones = create_buffer<T>(n_dofs, blas);
            blas->values(n_dofs, 1, ones->data());
            init(ones->data());
        }
        void init(const T *const guess_eigenvector) { /* :163-174 */
            temp_ = create_buffer<T>(n_dofs, blas);
            p_    = create_buffer<T>(n_dofs, blas);
            blas->copy(n_dofs, guess_eigenvector, temp_->data());
            PowerMethod<T> pm;
            pm.blas                 = blas;
            eig_max                 = pm.max_eigen_value(*apply_op, eigen_solver_max_it, eigen_solver_tol, n_dofs, temp_->data(),
                                         p_->data());
            power_method_iterations = pm.iterations;
        }

        int apply(const T *const b, T *const x) override { return precond_apply(b, x, p_->data(), temp_->data()); }

        /* :183-242 */
        int precond_apply(const T *const rhs, T *const x, T *const p, T *const temp) {
            if (!good()) return B200FEM_FAILURE;
            if (max_it <= 0) {
                iterations_ = 0;
                return 0;
            }
            const ptrdiff_t n        = n_dofs;
            const T         emax     = this->eig_max * scale_eig_max;
            const T         emin     = scale_eig_min * emax;
            const T         eig_avg  = (emin + emax) / 2;
            const T         eig_diff = (emin - emax) / 2;

            T alpha = T(1) / eig_avg;
            T beta  = 0;
            T dea   = 0;

            if (is_initial_guess_zero) {
                blas->zaxpby(n, T(-1), rhs, T(0), rhs, p);
            } else {
                if (apply_into(*apply_op, *blas, x, p)) return B200FEM_FAILURE;
                blas->axpy(n, T(-1), rhs, p);
            }
            blas->axpy(n, -alpha, p, x);
            iterations_ = 1;
            if (max_it == 1) return 0;

            dea   = eig_diff * alpha;
            beta  = T(0.5) * dea * dea;
            alpha = T(1) / (eig_avg - (beta / alpha));
            if (step(n, alpha, beta, rhs, x, p, temp)) return B200FEM_FAILURE;
            iterations_ = 2;
            if (max_it == 2) return 0;

            for (; iterations_ < max_it; iterations_++) {
                dea   = eig_diff * alpha;
                beta  = T(0.25) * dea * dea;
                alpha = T(1) / (eig_avg - (beta / alpha));
                if (step(n, alpha, beta, rhs, x, p, temp)) return B200FEM_FAILURE;
            }
            return blas->good() ? 0 : B200FEM_FAILURE;
        }

    private:
        int step(const ptrdiff_t n, const T alpha, const T beta, const T *rhs, T *x, T *p, T *temp) {
            if (apply_into(*apply_op, *blas, x, temp)) return B200FEM_FAILURE;
            blas->cheb3_iteration_update(n, alpha, beta, rhs, temp, p, x);
            return B200FEM_SUCCESS;
        }
    };

    /* ---------------------------------------------------------------------------------------------------- */
    /* Multigrid (level 0 is the finest)                                                                    */
    /* ---------------------------------------------------------------------------------------------------- */
    template <typename T>
    class Multigrid final : public Operator<T> {
    public:
        enum CycleType { V_CYCLE = 1, W_CYCLE = 2 };
        enum CycleReturnCode { CYCLE_CONTINUE = 0, CYCLE_CONVERGED = 1, CYCLE_FAILURE = 2 };

        std::shared_ptr<BLAS<T>> blas;
        bool                     verbose{false};
        /* fine-level residual norm at every cycle (what the reference prints, sfem_Multigrid.hpp:274-296) */
        std::vector<double> residual_history;
        int                 coarse_failures{0};

        void add_level(const std::shared_ptr<Operator<T>> &op, const std::shared_ptr<Operator<T>> &smoother_or_solver,
                       const std::shared_ptr<Operator<T>> &prolongation, const std::shared_ptr<Operator<T>> &restriction) {
            operator_.push_back(op);
            smoother_.push_back(smoother_or_solver);
            prolongation_.push_back(prolongation);
            restriction_.push_back(restriction);
        }
        void set_max_it(const int v) { max_it_ = v; }
        void set_atol(const T v) { atol_ = v; }
        void set_cycle_type(const int v) { cycle_type_ = v; }
        int  n_levels() const { return (int)smoother_.size(); }
        int  iterations() const { return iterations_; }
        std::ptrdiff_t rows() const override { return smoother_[0]->rows(); }
        std::ptrdiff_t cols() const override { return smoother_[0]->cols(); }

        int apply(const T *const rhs, T *const x) override {
            if (mem_.empty()) init();
            mem_[0].solution = x;
            mem_[0].rhs      = (T *)rhs;
            residual_history.clear();
            for (iterations_ = 0; iterations_ < max_it_; iterations_++) {
                const CycleReturnCode ret = cycle(0);
                if (ret == CYCLE_CONVERGED) break;
                if (ret == CYCLE_FAILURE) return B200FEM_FAILURE;
            }
            return blas->good() ? B200FEM_SUCCESS : B200FEM_FAILURE;
        }

    private:
        struct Memory {
            T              *rhs{nullptr}, *solution{nullptr}, *work{nullptr};
            ptrdiff_t       n{0};
            SharedBuffer<T> own_rhs, own_solution, own_work;
        };
        std::vector<std::shared_ptr<Operator<T>>> operator_, smoother_, prolongation_, restriction_;
        std::vector<Memory>                       mem_;
        int                                       max_it_{10};
        int                                       iterations_{0};
        int                                       cycle_type_{V_CYCLE};
        T                                         atol_{1e-10};
        T                                         norm_residual_0{1}, norm_residual_previous{1};

        void init() {
            mem_.resize(n_levels());
            for (int l = 0; l < n_levels(); l++) {
                auto &m = mem_[l];
                m.n     = smoother_[l]->rows();
                if (l != 0) {
                    m.own_solution = create_buffer<T>(m.n, blas);
                    m.own_rhs      = create_buffer<T>(m.n, blas);
                    m.solution     = m.own_solution->data();
                    m.rhs          = m.own_rhs->data();
                }
                m.own_work = create_buffer<T>(m.n, blas);
                m.work     = m.own_work->data();
            }
        }

        /* sfem_Multigrid.hpp:242-346 */
        CycleReturnCode cycle(const int level) {
            auto &mem      = mem_[level];
            auto &smoother = smoother_[level];
            if (level == n_levels() - 1) {
                blas->zeros(mem.n, mem.solution);
                /* a coarse solver that stops at max_it returns failure; the reference only asserts on it
                 * (sfem_Multigrid.hpp:311), i.e. a release build carries on with the best coarse correction */
                if (smoother->apply(mem.rhs, mem.solution)) coarse_failures++;
                return blas->good() ? CYCLE_CONTINUE : CYCLE_FAILURE;
            }
            auto &op           = operator_[level];
            auto &restriction  = restriction_[level];
            auto &prolongation = prolongation_[level + 1];
            auto &coarse       = mem_[level + 1];

            for (int k = 0; k < cycle_type_; k++) {
                if (smoother->apply(mem.rhs, mem.solution)) return CYCLE_FAILURE;

                if (apply_into(*op, *blas, mem.solution, mem.work)) return CYCLE_FAILURE;
                blas->axpby(mem.n, 1, mem.rhs, -1, mem.work);
                if (level == 0) {
                    const T norm_residual = blas->norm2(mem.n, mem.work);
                    residual_history.push_back((double)norm_residual);
                    if (iterations_ == 0) {
                        norm_residual_0 = norm_residual;
                        if (verbose) std::printf("Multigrid\niter\tabs\t\trel\t\trate\n%d\t%g\t-\t\t-\n", iterations_, (double)norm_residual);
                    } else if (verbose) {
                        std::printf("%d\t%g\t%g\t%g\n", iterations_, (double)norm_residual,
                                    (double)(norm_residual / norm_residual_0), (double)(norm_residual / norm_residual_previous));
                    }
                    norm_residual_previous = norm_residual;
                    if (norm_residual < atol_) return CYCLE_CONVERGED;
                }

                if (apply_into(*restriction, *blas, mem.work, coarse.rhs)) return CYCLE_FAILURE;
                blas->zeros(coarse.n, coarse.solution);

                if (cycle(level + 1) == CYCLE_FAILURE) return CYCLE_FAILURE;

                if (apply_into(*prolongation, *blas, coarse.solution, mem.work)) return CYCLE_FAILURE;
                blas->axpby(mem.n, 1, mem.work, 1, mem.solution);

                if (smoother->apply(mem.rhs, mem.solution)) return CYCLE_FAILURE;
            }
            return CYCLE_CONTINUE;
        }
    };

    /* ---------------------------------------------------------------------------------------------------- */
    /* sshex8 level transfer as Operator<T> (create_hierarchical_restriction / _prolongation,               */
    /* frontend/sfem_API.hpp:503-584), with the zero-constraints wrapper of mgsolve.cpp:232-241             */
    /* ---------------------------------------------------------------------------------------------------- */
    template <typename T>
    std::shared_ptr<Operator<T>> make_restriction(b200fem_transfer_t t, int vec_size,
                                                  const std::shared_ptr<Function<T>> &coarse_constraints = nullptr,
                                                  void                               *stream             = nullptr) {
        const ptrdiff_t rows = b200fem_transfer_coarse_nodes(t) * vec_size, cols = b200fem_transfer_fine_nodes(t) * vec_size;
        return make_op<T>(rows, cols, [=](const T *const from, T *const to) {
            B200FEM_TRY(b200fem_transfer_restrict(t, RealType<T>::value, vec_size, from, to, stream));
            return coarse_constraints ? coarse_constraints->apply_zero_constraints(to) : B200FEM_SUCCESS;
        });
    }
    template <typename T>
    std::shared_ptr<Operator<T>> make_prolongation(b200fem_transfer_t t, int vec_size,
                                                   const std::shared_ptr<Function<T>> &fine_constraints = nullptr,
                                                   void                               *stream           = nullptr) {
        const ptrdiff_t rows = b200fem_transfer_fine_nodes(t) * vec_size, cols = b200fem_transfer_coarse_nodes(t) * vec_size;
        auto            fn   = [=](const T *const from, T *const to) {
            B200FEM_TRY(b200fem_transfer_prolongate(t, RealType<T>::value, vec_size, from, to, stream));
            return fine_constraints ? fine_constraints->apply_zero_constraints(to) : B200FEM_SUCCESS;
        };
        /* the prolongation kernel writes every fine node: the same callable serves as the overwriting form */
        return make_op<T>(rows, cols, fn, fn);
    }

    /* ---------------------------------------------------------------------------------------------------- */
    /* Distributed layer: one process per GPU                                                               */
    /* ---------------------------------------------------------------------------------------------------- */
    /* smesh::Communicator as far as the path uses it: rank / size / sum. Owns the NCCL communicator. */
    class Communicator {
    public:
        /* `unique_id`: B200FEM_UNIQUE_ID_BYTES from b200fem_comm_unique_id on rank 0, handed round by the launcher */
        static std::shared_ptr<Communicator> create(int world, int rank, const void *unique_id) {
            b200fem_comm_t c = nullptr;
            if (b200fem_comm_create(&c, world, rank, unique_id)) {
                std::fprintf(stderr, "[b200fem] Communicator::create failed: %s\n", b200fem_last_error());
                return nullptr;
            }
            return std::shared_ptr<Communicator>(new Communicator(c));
        }
        ~Communicator() { b200fem_comm_destroy(c_); }
        int            rank() const { return b200fem_comm_rank(c_); }
        int            size() const { return b200fem_comm_size(c_); }
        b200fem_comm_t get() const { return c_; }
        double         sum(double v) const {
            b200fem_comm_allreduce_host(c_, &v, 1, B200FEM_REDUCE_SUM, nullptr);
            return v;
        }
        double max(double v) const {
            b200fem_comm_allreduce_host(c_, &v, 1, B200FEM_REDUCE_MAX, nullptr);
            return v;
        }
        void barrier() const { b200fem_comm_barrier(c_, nullptr); }

    private:
        explicit Communicator(b200fem_comm_t c) : c_(c) {}
        b200fem_comm_t c_;
    };

    /* algebra/parallel/sfem_ParallelOperator.hpp:13-25: rows() / cols() count OWNED dofs, vectors are allocated
     * with *_allocation_size() entries (owned + ghosts) */
    template <typename T>
    class ParallelOperator : public Operator<T> {
    public:
        virtual std::shared_ptr<Communicator> comm() const                = 0;
        virtual std::ptrdiff_t                row_allocation_size() const = 0;
        virtual std::ptrdiff_t                col_allocation_size() const = 0;
    };

    /* Halo plan of one space (smesh::Exchange): owns the b200fem_halo_t */
    class HaloExchange {
    public:
        HaloExchange(b200fem_halo_t h, std::shared_ptr<Communicator> comm) : h_(h), comm_(std::move(comm)) {}
        ~HaloExchange() { b200fem_halo_destroy(h_); }
        HaloExchange(const HaloExchange &)            = delete;
        HaloExchange &operator=(const HaloExchange &) = delete;
        int  gather(void *x_local, void *stream = nullptr) const { return b200fem_halo_gather(h_, x_local, stream); }
        int  gather_begin(void *x_local, void *stream = nullptr) const { return b200fem_halo_gather_begin(h_, x_local, stream); }
        int  gather_end(void *x_local, void *stream = nullptr) const { return b200fem_halo_gather_end(h_, x_local, stream); }
        int  gather_deferred_begin(void *x_local, void *stream, const unsigned long long **flag, unsigned long long *want) const {
            return b200fem_halo_gather_deferred_begin(h_, x_local, stream, flag, want);
        }
        int  gather_deferred_launch(void *x_local) const { return b200fem_halo_gather_deferred_launch(h_, x_local); }
        b200fem_halo_t                get() const { return h_; }
        std::shared_ptr<Communicator> comm() const { return comm_; }

    private:
        b200fem_halo_t                h_;
        std::shared_ptr<Communicator> comm_;
    };

    /* y(owned rows) = A x: gather the ghost block of x from the owners, then the local operator over the owned + aura
     * elements; interface rows are completed by the redundant aura computation, there is no reverse scatter
     * (frontend/sfem_ParallelMatrixFreeOperator.cpp:78-105). The reference stages x and y through scratch vectors of
     * local size; here the caller's vectors already have col_/row_allocation_size() entries (DeviceBLAS::
     * reserve_allocation), so the exchange and the kernels work in place and the ghost block of x is scratch that
     * every apply overwrites. Rows >= rows() of y hold partial sums and are never read. `in_halo` exchanges the
     * input space (it differs from the output space for the level transfers). */
    template <typename T>
    class ParallelMatrixFreeOperator final : public ParallelOperator<T> {
    public:
        ParallelMatrixFreeOperator(std::shared_ptr<Operator<T>> local, std::shared_ptr<HaloExchange> in_halo,
                                   std::ptrdiff_t owned_rows, std::ptrdiff_t owned_cols, void *stream = nullptr)
            : local_(std::move(local)), halo_(std::move(in_halo)), rows_(owned_rows), cols_(owned_cols), stream_(stream) {}

        int apply(const T *const x, T *const y) override {
            B200FEM_TRY(halo_->gather(const_cast<T *>(x), stream_));
            return local_->apply(x, y);
        }
        /* local operators without an overwriting form (the restriction) get their whole local output zeroed here:
         * the solvers only zero the owned rows, and the ghost rows would otherwise pile up partial sums */
        bool can_overwrite() const override { return true; }
        /* Elements [e0, e1) of the local operator read no ghost value: apply_overwrite then runs them between the two
         * halves of the gather (owner side first, receiver side after), and the ghost-dependent elements [0, e0) and
         * [e1, ne) last -- everything on one stream, so the wait for the peers sits behind ~97 % of the apply instead of in
         * front of it. Ignored when the local operator cannot be applied in pieces. */
        void set_independent_range(std::ptrdiff_t e0, std::ptrdiff_t e1) {
            ind0_ = e0;
            ind1_ = e1;
        }
        bool in_kernel_wait() const { return ind1_ > ind0_ && ind0_ == 0 && local_->can_wait_in_kernel(); }
        bool overlapped() const { return in_kernel_wait() || (split_launches && ind1_ > ind0_ && local_->range_elements() >= ind1_); }
        /* without the in-kernel wait: the pieces as separate launches on one stream (measured slower than gather-then-apply
         * on 2 GPUs at C3: the extra persistent launch costs more than the wait it hides; opt-in) */
        bool split_launches = false;
        int  apply_overwrite(const T *const x, T *const y) override {
            if (in_kernel_wait()) {
                /* one launch: the gather (push over NVLink, wait for the peers, copy into the ghost block) runs on the halo's
                 * own stream and one SM while the kernel already works on the elements before ind1_; its last groups wait
                 * for the ghost block inside the kernel */
                const unsigned long long *flag = nullptr;
                unsigned long long        want = 0;
                B200FEM_TRY(halo_->gather_deferred_begin(const_cast<T *>(x), stream_, &flag, &want));
                if (!flag) return local_->apply_overwrite(x, y); /* the gather went onto this stream in the ordinary way */
                B200FEM_TRY(halo_->gather_deferred_launch(const_cast<T *>(x)));
                return local_->apply_overwrite_after(x, y, ind1_, flag, want);
            }
            if (overlapped()) {
                const std::ptrdiff_t ne = local_->range_elements();
                B200FEM_TRY(halo_->gather_begin(const_cast<T *>(x), stream_));
                B200FEM_TRY(local_->apply_overwrite_range(x, y, ind0_, ind1_, true, false));
                B200FEM_TRY(halo_->gather_end(const_cast<T *>(x), stream_));
                if (ind0_ > 0) B200FEM_TRY(local_->apply_overwrite_range(x, y, 0, ind0_, false, false));
                return local_->apply_overwrite_range(x, y, ind1_, ne, false, true);  /* an empty range still finishes the apply */
            }
            B200FEM_TRY(halo_->gather(const_cast<T *>(x), stream_));
            if (local_->can_overwrite()) return local_->apply_overwrite(x, y);
            B200FEM_TRY(b200fem_blas_zeros(RealType<T>::value, local_->rows(), y, stream_));
            return local_->apply(x, y);
        }
        std::ptrdiff_t                rows() const override { return rows_; }
        std::ptrdiff_t                cols() const override { return cols_; }
        std::ptrdiff_t                row_allocation_size() const override { return local_->rows(); }
        std::ptrdiff_t                col_allocation_size() const override { return local_->cols(); }
        std::shared_ptr<Communicator> comm() const override { return halo_->comm(); }
        void reserve(DeviceBLAS<T> &blas) const {
            blas.reserve_allocation((std::size_t)rows_, (std::size_t)local_->rows());
            blas.reserve_allocation((std::size_t)cols_, (std::size_t)local_->cols());
        }

    private:
        std::shared_ptr<Operator<T>>  local_;
        std::shared_ptr<HaloExchange> halo_;
        std::ptrdiff_t                rows_, cols_, ind0_ = 0, ind1_ = 0;
        void                         *stream_;
    };

    /* algebra/parallel/sfem_parallel_cg.hpp:257-400 is ConjugateGradient with every reduction followed by comm->sum and
     * vectors of row_allocation_size(): the same recurrences and stopping rule, so it is the same class here, bound to a
     * ParallelOperator and a DeviceBLAS whose reductions run over the communicator. */
    template <typename T>
    std::shared_ptr<ConjugateGradient<T>> make_parallel_cg(const std::shared_ptr<ParallelOperator<T>> &op,
                                                           const std::shared_ptr<DeviceBLAS<T>>       &blas) {
        auto cg = std::make_shared<ConjugateGradient<T>>();
        blas->comm = op->comm() ? op->comm()->get() : nullptr;
        blas->reserve_allocation((std::size_t)op->rows(), (std::size_t)op->row_allocation_size());
        cg->blas = blas;
        cg->set_op(op);
        return cg;
    }

}  // namespace b200fem
#endif
