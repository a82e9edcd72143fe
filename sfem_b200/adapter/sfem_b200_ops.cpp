/* See sfem_b200_ops.hpp. Pattern: GPULaplacian / GPULinearElasticity of the reference
 * (frontend/cuda/sfem_Function_incore_cuda.cpp:545-750, 753-1100), with the plans of libb200fem.so in place of the
 * reference's device caches. */
#include "sfem_b200_ops.hpp"

#include <cstdlib>

#include "smesh_mesh.hpp"
#include "smesh_semistructured.hpp"

namespace sfem {

    namespace {
        constexpr int kRealType = sizeof(real_t) == 8 ? B200FEM_F64 : B200FEM_F32;

        /* VTK corner -> (x, y, z) in {0,1}^3: the 8 macro corners of an sshex8 element are the lattice rows
         * sshex8_lidx(L, cx L, cy L, cz L) (operators/sshex8/sshex8.hpp:18) */
        const int kHex8Corners[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};

        struct DomainGeometry {
            int                  type{0}, level{0};
            ptrdiff_t            ne{0}, nn{0};
            idx_t              **elements{nullptr};
            std::vector<idx_t *> corners; /* rows of the 8 (macro) corners for the geometry fills */
        };

        DomainGeometry geometry_of(const std::shared_ptr<FunctionSpace> &space, const OpDomain &d) {
            DomainGeometry g;
            auto           mesh = space->mesh_ptr();
            g.ne                = d.block->n_elements();
            g.nn                = mesh->n_nodes();
            g.elements          = d.block->elements()->data();
            if (is_semistructured_type(d.element_type)) {
                g.type  = B200FEM_SSHEX8;
                g.level = smesh::semistructured_level(*mesh);
                const int N = g.level + 1;
                for (auto &c : kHex8Corners) g.corners.push_back(g.elements[(c[2] * g.level * N + c[1] * g.level) * N + c[0] * g.level]);
            } else if (d.element_type == smesh::HEX8) {
                g.type = B200FEM_HEX8;
                for (int v = 0; v < 8; v++) g.corners.push_back(g.elements[v]);
            } else if (d.element_type == smesh::TET4 || d.element_type == smesh::TET10 || d.element_type == smesh::MACRO_TET4) {
                /* 10-node tetrahedra take the affine geometry of their 4 corners (rows 0-3), like
                 * laplacian_apply_opt / linear_elasticity_apply_opt do (operators/laplacian.cpp:221-226) */
                g.type = d.element_type == smesh::TET4 ? B200FEM_TET4 : d.element_type == smesh::TET10 ? B200FEM_TET10 : B200FEM_MACRO_TET4;
                for (int v = 0; v < 4; v++) g.corners.push_back(g.elements[v]);
            }
            return g;
        }
    }  // namespace

    /* ------------------------------------------------------------------------------------------ B200Op */
    int B200Op::initialize_domains(const std::vector<std::string> &block_names) {
        domains = std::make_shared<MultiDomainOp>(space, block_names);
        for (auto &n2d : domains->domains()) {
            OpDomain    &domain = n2d.second;
            b200fem_op_t plan   = create_plan(domain);
            if (!plan) {
                fprintf(stderr, "[b200fem] %s: %s\n", name(), b200fem_last_error());
                return SFEM_FAILURE;
            }
            domain.user_data = std::shared_ptr<void>((void *)plan, [](void *p) { b200fem_op_destroy((b200fem_op_t)p); });
        }
        return SFEM_SUCCESS;
    }

    int B200Op::apply(const real_t *const, const real_t *const h, real_t *const out) {
        return domains->iterate(
                [&](const OpDomain &d) { return b200fem_op_apply((b200fem_op_t)d.user_data.get(), h, out, stream); });
    }

    int B200Op::hessian_diag(const real_t *const, real_t *const values) {
        return domains->iterate(
                [&](const OpDomain &d) { return b200fem_op_diag((b200fem_op_t)d.user_data.get(), values, stream); });
    }

    std::vector<std::string> B200Op::selected_block_names() const {
        std::vector<std::string> ret;
        if (domains)
            for (const auto &it : domains->domains()) ret.push_back(it.first);
        return ret;
    }

    /* ------------------------------------------------------------------------------------ B200Laplacian */
    std::unique_ptr<Op> B200Laplacian::create(const std::shared_ptr<FunctionSpace> &space) {
        assert(1 == space->block_size());
        return std::make_unique<B200Laplacian>(space);
    }

    b200fem_op_t B200Laplacian::create_plan(const OpDomain &domain) {
        DomainGeometry g = geometry_of(space, domain);
        if (!g.type) return nullptr;
        /* the FFF cache of Laplacian::initialize (smesh::FFF::create_AoS, frontend/ops/sfem_Laplacian.cpp:88): fp32,
         * [ne * 6]; for sshex8 one set per macro element from its 8 corners */
        std::vector<float> fff((size_t)g.ne * 6);
        const int          geo_type = g.corners.size() == 4 ? B200FEM_TET4 : B200FEM_HEX8;
        if (b200fem_fff_fill(geo_type, g.ne, g.nn, g.corners.data(), space->mesh_ptr()->points()->data(), fff.data())) return nullptr;
        b200fem_op_t plan = nullptr;
        if (b200fem_laplacian_create(&plan, g.type, g.level, g.ne, g.nn, g.elements, fff.data(), kRealType)) return nullptr;
        return plan;
    }

    std::shared_ptr<Op> B200Laplacian::derefine_op(const std::shared_ptr<FunctionSpace> &derefined_space) {
        /* pattern of GPULaplacian::derefine_op (sfem_Function_incore_cuda.cpp:600-629): a new operator on the coarse space */
        auto ret = std::make_shared<B200Laplacian>(derefined_space);
        if (ret->initialize(selected_block_names()) != SFEM_SUCCESS) return nullptr;
        return ret;
    }

    std::shared_ptr<Op> B200Laplacian::clone() const {
        auto ret     = std::make_shared<B200Laplacian>(space);
        ret->domains = domains; /* plans are immutable and shared */
        ret->stream  = stream;
        return ret;
    }

    /* ----------------------------------------------------------------------------- B200LinearElasticity */
    std::unique_ptr<Op> B200LinearElasticity::create(const std::shared_ptr<FunctionSpace> &space) {
        assert(space->mesh_ptr()->spatial_dimension() == space->block_size());
        return std::make_unique<B200LinearElasticity>(space);
    }

    int B200LinearElasticity::initialize(const std::vector<std::string> &block_names) {
        /* material defaults and environment overrides of LinearElasticity::initialize
         * (frontend/ops/sfem_LinearElasticity.cpp:256-260) */
        real_t SFEM_SHEAR_MODULUS        = mu;
        real_t SFEM_FIRST_LAME_PARAMETER = lambda;
        SFEM_READ_ENV(SFEM_SHEAR_MODULUS, atof);
        SFEM_READ_ENV(SFEM_FIRST_LAME_PARAMETER, atof);
        mu     = SFEM_SHEAR_MODULUS;
        lambda = SFEM_FIRST_LAME_PARAMETER;
        return initialize_domains(block_names);
    }

    b200fem_op_t B200LinearElasticity::create_plan(const OpDomain &domain) {
        DomainGeometry g = geometry_of(space, domain);
        if (!g.type) {
            fprintf(stderr, "[b200fem] linear elasticity: tet4 / tet10 / macro_tet4 / hex8 / sshex8 blocks only\n");
            return nullptr;
        }
        /* JacobianAdjugateAndDeterminant cache (sfem_LinearElasticity.cpp:275-279): adjugate [ne * 9] and determinant [ne],
         * fp32, at the reference point (1/2, 1/2, 1/2) of the (macro) element */
        std::vector<float> adj((size_t)g.ne * 9), det((size_t)g.ne);
        if (b200fem_adjugate_and_det_fill(g.corners.size() == 4 ? B200FEM_TET4 : B200FEM_HEX8, g.ne, g.nn, g.corners.data(), space->mesh_ptr()->points()->data(), adj.data(),
                                          det.data()))
            return nullptr;
        const real_t m = domain.parameters ? domain.parameters->get_real_value("mu", mu) : mu;
        const real_t l = domain.parameters ? domain.parameters->get_real_value("lambda", lambda) : lambda;
        b200fem_op_t plan = nullptr;
        if (b200fem_linear_elasticity_create(&plan, g.type, g.level, g.ne, g.nn, g.elements, adj.data(), det.data(), (double)m, (double)l,
                                             kRealType))
            return nullptr;
        return plan;
    }

    int B200LinearElasticity::hessian_block_diag_sym(const real_t *const, real_t *const values) {
        return domains->iterate(
                [&](const OpDomain &d) { return b200fem_op_block_diag_sym((b200fem_op_t)d.user_data.get(), values, stream); });
    }

    std::shared_ptr<Op> B200LinearElasticity::derefine_op(const std::shared_ptr<FunctionSpace> &derefined_space) {
        auto ret    = std::make_shared<B200LinearElasticity>(derefined_space);
        ret->mu     = mu;
        ret->lambda = lambda;
        if (ret->initialize(selected_block_names()) != SFEM_SUCCESS) return nullptr;
        return ret;
    }

    std::shared_ptr<Op> B200LinearElasticity::clone() const {
        auto ret     = std::make_shared<B200LinearElasticity>(space);
        ret->domains = domains;
        ret->mu      = mu;
        ret->lambda  = lambda;
        ret->stream  = stream;
        return ret;
    }

    /* ------------------------------------------------------------------------------ B200VectorLaplacian */
    std::unique_ptr<Op> B200VectorLaplacian::create(const std::shared_ptr<FunctionSpace> &space) {
        return std::make_unique<B200VectorLaplacian>(space);
    }

    b200fem_op_t B200VectorLaplacian::create_plan(const OpDomain &domain) {
        DomainGeometry g = geometry_of(space, domain);
        if (g.type != B200FEM_HEX8 && g.type != B200FEM_SSHEX8) {
            fprintf(stderr, "[b200fem] vector laplacian: hex8 and sshex8 blocks only\n");
            return nullptr;
        }
        std::vector<float> fff((size_t)g.ne * 6);
        if (b200fem_fff_fill(B200FEM_HEX8, g.ne, g.nn, g.corners.data(), space->mesh_ptr()->points()->data(), fff.data())) return nullptr;
        b200fem_op_t plan = nullptr;
        const int    vs   = space->block_size();
        const int    rc   = g.type == B200FEM_SSHEX8
                                    ? b200fem_sshex8_vector_laplacian_create(&plan, g.level, g.ne, g.nn, g.elements, fff.data(), vs, kRealType)
                                    : b200fem_vector_laplacian_create(&plan, g.type, g.ne, g.nn, g.elements, fff.data(), vs, kRealType);
        return rc ? nullptr : plan;
    }

    std::shared_ptr<Op> B200VectorLaplacian::derefine_op(const std::shared_ptr<FunctionSpace> &derefined_space) {
        auto ret = std::make_shared<B200VectorLaplacian>(derefined_space);
        if (ret->initialize(selected_block_names()) != SFEM_SUCCESS) return nullptr;
        return ret;
    }

    std::shared_ptr<Op> B200VectorLaplacian::clone() const {
        auto ret     = std::make_shared<B200VectorLaplacian>(space);
        ret->domains = domains;
        ret->stream  = stream;
        return ret;
    }

    /* ------------------------------------------------------------------------------------------ B200Mass */
    std::unique_ptr<Op> B200Mass::create(const std::shared_ptr<FunctionSpace> &space) {
        assert(1 == space->block_size());
        return std::make_unique<B200Mass>(space);
    }

    b200fem_op_t B200Mass::create_plan(const OpDomain &domain) {
        DomainGeometry g = geometry_of(space, domain);
        if (g.type != B200FEM_HEX8 && g.type != B200FEM_TET4 && g.type != B200FEM_TET10 && g.type != B200FEM_SSHEX8) {
            fprintf(stderr, "[b200fem] mass: hex8 (apply + lumped), tet4, tet10 and sshex8 (lumped) blocks only\n");
            return nullptr;
        }
        b200fem_op_t plan = nullptr; /* apply_mass / assemble_lumped_mass read the mesh points, no cached geometry */
        if (g.type == B200FEM_SSHEX8) { /* assemble_lumped_mass -> affine_sshex8_mass_lumped (operators/mass.cpp:74-77) */
            if (b200fem_sshex8_lumped_mass_create(&plan, g.level, g.ne, g.nn, g.elements, space->mesh_ptr()->points()->data(), kRealType))
                return nullptr;
            return plan;
        }
        if (b200fem_mass_create(&plan, g.type, g.ne, g.nn, g.elements, space->mesh_ptr()->points()->data(), kRealType)) return nullptr;
        return plan;
    }

    std::shared_ptr<Op> B200Mass::clone() const {
        auto ret     = std::make_shared<B200Mass>(space);
        ret->domains = domains;
        ret->stream  = stream;
        return ret;
    }

    void register_b200_ops() {
        Factory::register_op("gpu:Mass", &B200Mass::create);
        Factory::register_op("gpu:LumpedMass", &B200Mass::create);
        Factory::register_op("ss:gpu:LumpedMass", &B200Mass::create);
        Factory::register_op("gpu:VectorLaplacian", &B200VectorLaplacian::create);
        Factory::register_op("ss:gpu:VectorLaplacian", &B200VectorLaplacian::create);
        Factory::register_op("gpu:Laplacian", &B200Laplacian::create);
        Factory::register_op("ss:gpu:Laplacian", &B200Laplacian::create);
        Factory::register_op("gpu:LinearElasticity", &B200LinearElasticity::create);
        Factory::register_op("ss:gpu:LinearElasticity", &B200LinearElasticity::create);
    }

    template class B200BLAS<real_t>;

}  // namespace sfem
