/* sfem_b200_ops.hpp -- the file a maintainer adds on the sfem side (e.g. frontend/b200/) to plug libb200fem.so into the
 * reference: sfem::Op adapters for the four factory keys the reference's own CUDA twin owns
 * ("gpu:Laplacian", "ss:gpu:Laplacian", "gpu:LinearElasticity", "ss:gpu:LinearElasticity",
 * frontend/cuda/sfem_Function_incore_cuda.cpp:1811-1822) and a sfem::BLAS<T> over the vector entry points.
 *
 * Compiles against the reference's OWN headers (frontend/ops/sfem_Op.hpp, sfem_OpFactory.hpp, sfem_MultiDomainOp.hpp,
 * algebra/sfem_tpl_blas.hpp) and the C ABI of include/b200fem.h; nothing else from this repository.
 * In this repository it is built by oracle/Makefile (target `adapter`) against those headers where they lie under
 * /root/reference, with stand-ins for the absent smesh submodule (oracle/ref_shim/frontend_stub), and driven by
 * tests/test_gpu_adapter.py through sfem::create_op(space, "Laplacian", EXECUTION_SPACE_DEVICE). */
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "b200fem.h"
#include "sfem_MultiDomainOp.hpp"
#include "sfem_Op.hpp"
#include "sfem_OpFactory.hpp"
#include "sfem_tpl_blas.hpp"

namespace sfem {

    /* Registers the four keys in sfem::Factory; the last registration wins (frontend/ops/sfem_OpFactory.cpp:83-89), so
     * call it after register_device_ops() in sfem::initialize (frontend/sfem_context.cpp:11-17). */
    void register_b200_ops();

    /* common part of the two adapters: one b200fem_op_t per mesh block in OpDomain::user_data */
    class B200Op : public Op {
    public:
        std::shared_ptr<FunctionSpace> space;
        std::shared_ptr<MultiDomainOp> domains;
        void                          *stream{SFEM_DEFAULT_STREAM};
        smesh::ElemType                element_type{smesh::INVALID};

        explicit B200Op(const std::shared_ptr<FunctionSpace> &s) : space(s), element_type(s->element_type()) {}

        bool           is_linear() const override { return true; }
        ptrdiff_t      n_dofs_domain() const override { return space->n_dofs(); }
        ptrdiff_t      n_dofs_image() const override { return space->n_dofs(); }
        ExecutionSpace execution_space() const override { return EXECUTION_SPACE_DEVICE; }

        /* out += A h on DEVICE pointers (Op::apply, frontend/ops/sfem_Op.hpp:214) */
        int apply(const real_t *const /*x*/, const real_t *const h, real_t *const out) override;
        int gradient(const real_t *const x, real_t *const out) override { return apply(nullptr, x, out); }
        int hessian_diag(const real_t *const /*x*/, real_t *const values) override;
        int hessian_crs(const real_t *const, const count_t *const, const idx_t *const, real_t *const) override {
            SFEM_ERROR("b200fem operators are matrix-free: hessian_crs is not available\n");
            return SFEM_FAILURE;
        }
        int value(const real_t *, real_t *const) override {
            SFEM_ERROR("b200fem operators are matrix-free: value is not available\n");
            return SFEM_FAILURE;
        }
        void override_element_types(const std::vector<smesh::ElemType> &element_types) override {
            if (domains) domains->override_element_types(element_types);
        }
        void set_value_in_block(const std::string &block_name, const std::string &var_name, const real_t value) override {
            if (domains) domains->set_value_in_block(block_name, var_name, value);
        }
        std::vector<std::string> selected_block_names() const;

    protected:
        /* builds the plan of one domain; implemented by the two operators */
        virtual b200fem_op_t create_plan(const OpDomain &domain) = 0;
        int                  initialize_domains(const std::vector<std::string> &block_names);
    };

    class B200Laplacian final : public B200Op {
    public:
        using B200Op::B200Op;
        static std::unique_ptr<Op> create(const std::shared_ptr<FunctionSpace> &space);
        const char *name() const override { return is_semistructured_type(element_type) ? "ss:gpu:Laplacian" : "gpu:Laplacian"; }
        int         initialize(const std::vector<std::string> &block_names = {}) override { return initialize_domains(block_names); }
        std::shared_ptr<Op> derefine_op(const std::shared_ptr<FunctionSpace> &derefined_space) override;
        std::shared_ptr<Op> clone() const override;

    protected:
        b200fem_op_t create_plan(const OpDomain &domain) override;
    };

    class B200LinearElasticity final : public B200Op {
    public:
        real_t mu{1}, lambda{1};
        using B200Op::B200Op;
        static std::unique_ptr<Op> create(const std::shared_ptr<FunctionSpace> &space);
        const char *name() const override {
            return is_semistructured_type(element_type) ? "ss:gpu:LinearElasticity" : "gpu:LinearElasticity";
        }
        int initialize(const std::vector<std::string> &block_names = {}) override;
        int hessian_block_diag_sym(const real_t *const /*x*/, real_t *const values) override;
        std::shared_ptr<Op> derefine_op(const std::shared_ptr<FunctionSpace> &derefined_space) override;
        std::shared_ptr<Op> clone() const override;

    protected:
        b200fem_op_t create_plan(const OpDomain &domain) override;
    };

    /* VectorLaplacian (frontend/ops/sfem_VectorLaplacian.cpp): the scalar Laplacian on every component of an AoS field with
     * block_size components; hex8 and sshex8 blocks (vector_laplacian_apply_opt, operators/vector_laplacian.cpp:41-66) */
    class B200VectorLaplacian final : public B200Op {
    public:
        using B200Op::B200Op;
        static std::unique_ptr<Op> create(const std::shared_ptr<FunctionSpace> &space);
        const char *name() const override { return is_semistructured_type(element_type) ? "ss:gpu:VectorLaplacian" : "gpu:VectorLaplacian"; }
        int         initialize(const std::vector<std::string> &block_names = {}) override { return initialize_domains(block_names); }
        std::shared_ptr<Op> derefine_op(const std::shared_ptr<FunctionSpace> &derefined_space) override;
        std::shared_ptr<Op> clone() const override;

    protected:
        b200fem_op_t create_plan(const OpDomain &domain) override;
    };

    /* Mass / LumpedMass (frontend/ops/sfem_Mass.cpp, sfem_LumpedMass.cpp): apply = Mass::apply (hex8 blocks: apply_mass has no
     * other 3-D case), hessian_diag = LumpedMass::hessian_diag (hex8, tet4, sshex8: "ss:gpu:LumpedMass"). One class under all keys. */
    class B200Mass final : public B200Op {
    public:
        using B200Op::B200Op;
        static std::unique_ptr<Op> create(const std::shared_ptr<FunctionSpace> &space);
        const char *name() const override { return is_semistructured_type(element_type) ? "ss:gpu:LumpedMass" : "gpu:Mass"; }
        int         initialize(const std::vector<std::string> &block_names = {}) override { return initialize_domains(block_names); }
        std::shared_ptr<Op> clone() const override;

    protected:
        b200fem_op_t create_plan(const OpDomain &domain) override;
    };

    /* sfem::BLAS<T> over b200fem_blas_*; injected like CUDA_BLAS (algebra/cuda/sfem_cuda_solver.hpp:15-49):
     *     cg->blas = std::make_shared<B200BLAS<real_t>>(); cg->execution_space_ = EXECUTION_SPACE_DEVICE; */
    template <typename T>
    class B200BLAS final : public BLAS<T> {
        static constexpr int RT = (int)sizeof(T); /* B200FEM_F64 == 8, B200FEM_F32 == 4 */

    public:
        void *stream{nullptr};
        T    *allocate(const std::size_t n) override {
            void *p = nullptr;
            return b200fem_malloc(&p, (n ? n : 1) * sizeof(T)) == B200FEM_SUCCESS ? (T *)p : nullptr; /* zero-filled */
        }
        void destroy(void *a) override { b200fem_free(a); }
        void zeros(const std::size_t n, T *const x) override { ok_ &= !b200fem_blas_zeros(RT, (ptrdiff_t)n, x, stream); }
        void values(const std::size_t n, const T v, T *const x) override { ok_ &= !b200fem_blas_values(RT, (ptrdiff_t)n, v, x, stream); }
        void copy(const ptrdiff_t n, const T *const s, T *const d) override { ok_ &= !b200fem_blas_copy(RT, n, s, d, stream); }
        T    dot(const ptrdiff_t n, const T *const l, const T *const r) override {
            double v = 0;
            ok_ &= !b200fem_blas_dot(RT, n, l, r, &v, stream);
            return (T)v;
        }
        T norm2(const ptrdiff_t n, const T *const x) override {
            double v = 0;
            ok_ &= !b200fem_blas_norm2(RT, n, x, &v, stream);
            return (T)v;
        }
        void axpy(const ptrdiff_t n, const T a, const T *const x, T *const y) override { ok_ &= !b200fem_blas_axpy(RT, n, a, x, y, stream); }
        void axpby(const ptrdiff_t n, const T a, const T *const x, const T b, T *const y) override {
            ok_ &= !b200fem_blas_axpby(RT, n, a, x, b, y, stream);
        }
        void zaxpby(const ptrdiff_t n, const T a, const T *const x, const T b, const T *const y, T *const z) override {
            ok_ &= !b200fem_blas_zaxpby(RT, n, a, x, b, y, z, stream);
        }
        void scal(const std::ptrdiff_t n, const T a, T *const x) override { ok_ &= !b200fem_blas_scal(RT, n, a, x, stream); }
        void reciprocal(const std::ptrdiff_t n, const T a, T *const x) override { ok_ &= !b200fem_blas_reciprocal(RT, n, a, x, stream); }
        void xypaz(const ptrdiff_t n, const T *const x, const T *const y, const T a, T *const z) override {
            ok_ &= !b200fem_blas_xypaz(RT, n, x, y, a, z, stream);
        }
        bool good() const override { return ok_; }

    private:
        bool ok_{true};
    };

}  // namespace sfem
