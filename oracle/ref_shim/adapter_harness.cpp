/* TEST INFRASTRUCTURE ONLY. Drives sfem_b200/adapter/sfem_b200_ops.cpp -- compiled against the reference's real
 * frontend/ops/sfem_Op.hpp, sfem_OpFactory.hpp, sfem_MultiDomainOp.hpp (+ .cpp) and frontend/sfem_Parameters.cpp -- the
 * way the reference's drivers do: sfem::create_op(space, "Laplacian", EXECUTION_SPACE_DEVICE) -> op->initialize() ->
 * op->apply(nullptr, h, out) (drivers/bench/bench_op.exe.cpp:33-55, frontend/sfem_Function.cpp:484-489).
 *
 * What cannot come from the reference is provided here, as small as possible:
 *   * sfem::FunctionSpace: the class is the reference's declaration (frontend/sfem_FunctionSpace.hpp); its definition
 *     lives in sfem_FunctionSpace.cpp, which needs smesh. The members the Op path calls are defined below over the
 *     stand-in smesh::Mesh (oracle/ref_shim/frontend_stub/smesh_mesh.hpp).
 *   * sfem::Factory: declaration from the reference; frontend/ops/sfem_OpFactory.cpp cannot be compiled (it includes every
 *     operator of the library), so register_op / create_op / create_op_gpu are restated from its lines 83-123: a name ->
 *     function map, "gpu:" prefix for device ops, "ss:" prefix for semi-structured spaces, last registration wins. */
#include <cstring>
#include <iostream>
#include <map>

#include "sfem_b200_ops.hpp"
#include "smesh_mesh.hpp"

extern "C" void sfem_abort();

namespace sfem {

    /* ---- FunctionSpace over the stand-in mesh ------------------------------------------------------------ */
    class FunctionSpace::Impl {
    public:
        std::shared_ptr<Mesh> mesh;
        int                   block_size{1};
        smesh::ElemType       element_type{smesh::INVALID};
    };
    FunctionSpace::FunctionSpace() : impl_(std::make_unique<Impl>()) {}
    FunctionSpace::FunctionSpace(const std::shared_ptr<Mesh> &mesh, const int block_size, const smesh::ElemType element_type)
        : impl_(std::make_unique<Impl>()) {
        impl_->mesh         = mesh;
        impl_->block_size   = block_size;
        impl_->element_type = element_type != smesh::INVALID ? element_type : mesh->block(0)->element_type();
    }
    FunctionSpace::~FunctionSpace() = default;
    std::shared_ptr<Mesh> FunctionSpace::mesh_ptr() const { return impl_->mesh; }
    Mesh                 &FunctionSpace::mesh() { return *impl_->mesh; }
    bool                  FunctionSpace::has_semi_structured_mesh() const { return impl_->mesh->level_ > 0; }
    int                   FunctionSpace::block_size() const { return impl_->block_size; }
    ptrdiff_t             FunctionSpace::n_dofs() const { return impl_->mesh->n_nodes() * impl_->block_size; }
    size_t                FunctionSpace::n_blocks() const { return impl_->mesh->n_blocks(); }
    smesh::ElemType       FunctionSpace::element_type(const int block) const {
        return block == 0 ? impl_->element_type : impl_->mesh->block((size_t)block)->element_type();
    }

    /* ---- Factory, restated from frontend/ops/sfem_OpFactory.cpp:83-123 ------------------------------------- */
    class Factory::Impl {
    public:
        std::map<std::string, FactoryFunction> name_to_create;
    };
    Factory::Factory() : impl_(std::make_unique<Impl>()) {}
    Factory::~Factory() = default;
    Factory &Factory::instance() {
        static Factory instance_;
        return instance_;
    }
    void Factory::private_register_op(const std::string &name, FactoryFunction f) { impl_->name_to_create[name] = f; }
    void Factory::register_op(const std::string &name, FactoryFunction f) { instance().private_register_op(name, f); }
    std::string d_op_str(const std::string &name) { return "gpu:" + name; }
    std::shared_ptr<Op> Factory::create_op_gpu(const std::shared_ptr<FunctionSpace> &space, const char *name) {
        return Factory::create_op(space, d_op_str(name).c_str());
    }
    std::shared_ptr<Op> Factory::create_op(const std::shared_ptr<FunctionSpace> &space, const char *name) {
        std::string m_name = name;
        if (space->has_semi_structured_mesh()) m_name = "ss:" + m_name;
        auto &ntc = instance().impl_->name_to_create;
        auto  it  = ntc.find(m_name);
        if (it == ntc.end()) {
            std::cerr << "Unable to find op " << m_name << "\n";
            return nullptr;
        }
        return it->second(space);
    }
}  // namespace sfem

namespace {
    struct SpaceBox {
        std::vector<std::vector<idx_t>>  rows;
        std::vector<idx_t *>             row_ptrs;
        std::vector<std::vector<geom_t>> coords;
        std::vector<geom_t *>            coord_ptrs;
        std::shared_ptr<sfem::Mesh>      mesh;
        std::shared_ptr<sfem::FunctionSpace> space;
    };
    struct OpBox {
        std::shared_ptr<sfem::Op> op;
    };
}  // namespace

extern "C" {

/* element_type: smesh::ElemType value (4 tet4, 8 hex8, 800 sshex8); level > 0 marks the mesh semi-structured */
void *adapter_space_create(int element_type, int level, int nxe, ptrdiff_t ne, ptrdiff_t nn, const idx_t *const *elements,
                           const geom_t *const *points, int block_size) {
    auto b = new SpaceBox();
    b->rows.resize((size_t)nxe);
    for (int v = 0; v < nxe; v++) {
        b->rows[(size_t)v].assign(elements[v], elements[v] + ne);
        b->row_ptrs.push_back(b->rows[(size_t)v].data());
    }
    b->coords.resize(3);
    for (int d = 0; d < 3; d++) {
        b->coords[(size_t)d].assign(points[d], points[d] + nn);
        b->coord_ptrs.push_back(b->coords[(size_t)d].data());
    }
    b->mesh           = std::make_shared<sfem::Mesh>();
    b->mesh->n_nodes_ = nn;
    b->mesh->level_   = level;
    b->mesh->points_  = sfem::Buffer<geom_t *>::wrap(3, b->coord_ptrs.data());
    b->mesh->blocks_.push_back(std::make_shared<sfem::Mesh::Block>("default", (smesh::ElemType)element_type,
                                                                   sfem::Buffer<idx_t *>::wrap((size_t)nxe, b->row_ptrs.data()), ne));
    b->space = sfem::FunctionSpace::create(b->mesh, block_size);
    return b;
}
void adapter_space_destroy(void *s) { delete (SpaceBox *)s; }

/* sfem::create_op(space, name, EXECUTION_SPACE_DEVICE) after register_b200_ops(), then Op::initialize() */
void *adapter_create_op(void *space, const char *name) {
    static bool registered = false;
    if (!registered) {
        sfem::register_b200_ops();
        registered = true;
    }
    auto op = sfem::create_op(((SpaceBox *)space)->space, name, sfem::EXECUTION_SPACE_DEVICE);
    if (!op) return nullptr;
    if (op->initialize() != SFEM_SUCCESS) return nullptr;
    auto b = new OpBox();
    b->op  = op;
    return b;
}
void        adapter_op_destroy(void *o) { delete (OpBox *)o; }
const char *adapter_op_name(void *o) { return ((OpBox *)o)->op->name(); }
int         adapter_op_execution_space(void *o) { return (int)((OpBox *)o)->op->execution_space(); }
ptrdiff_t   adapter_op_n_dofs(void *o) { return ((OpBox *)o)->op->n_dofs_domain(); }
int adapter_op_apply(void *o, const real_t *h_dev, real_t *out_dev) { return ((OpBox *)o)->op->apply(nullptr, h_dev, out_dev); }
int adapter_op_gradient(void *o, const real_t *x_dev, real_t *out_dev) { return ((OpBox *)o)->op->gradient(x_dev, out_dev); }
int adapter_op_hessian_diag(void *o, real_t *values_dev) { return ((OpBox *)o)->op->hessian_diag(nullptr, values_dev); }
int adapter_op_hessian_block_diag_sym(void *o, real_t *values_dev) { return ((OpBox *)o)->op->hessian_block_diag_sym(nullptr, values_dev); }
void *adapter_op_clone(void *o) {
    auto c = ((OpBox *)o)->op->clone();
    if (!c) return nullptr;
    auto b = new OpBox();
    b->op  = c;
    return b;
}
/* Op::derefine_op onto another space (multigrid coarse operators, frontend/cuda/sfem_Function_incore_cuda.cpp:600-629) */
void *adapter_op_derefine(void *o, void *coarse_space) {
    auto c = ((OpBox *)o)->op->derefine_op(((SpaceBox *)coarse_space)->space);
    if (!c) return nullptr;
    auto b = new OpBox();
    b->op  = c;
    return b;
}
/* sfem::BLAS<real_t> through the adapter's B200BLAS: y = alpha x + beta y, then (x, y) */
double adapter_blas_axpby_dot(ptrdiff_t n, double alpha, const real_t *x_dev, double beta, real_t *y_dev) {
    sfem::B200BLAS<real_t>              blas;
    std::shared_ptr<sfem::BLAS<real_t>> keep;  // the base-class interface is what the solvers hold
    sfem::BLAS<real_t>                 &b = blas;
    b.axpby(n, (real_t)alpha, x_dev, (real_t)beta, y_dev);
    return (double)b.dot(n, x_dev, y_dev);
}
}
