/* smesh::Mesh as far as the operator adapters read it: blocks (name, element type, SoA connectivity), points, counts. */
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "smesh_elem_type.hpp"
#include "smesh_types.hpp"

namespace smesh {
    typedef int block_idx_t;

    class Sideset {};

    class Mesh {
    public:
        class Block {
        public:
            Block(std::string name, ElemType type, SharedBuffer<idx_t *> elements, ptrdiff_t n_elements)
                : name_(std::move(name)), type_(type), elements_(std::move(elements)), n_elements_(n_elements) {}
            const std::string            &name() const { return name_; }
            ElemType                      element_type() const { return type_; }
            const SharedBuffer<idx_t *> &elements() const { return elements_; }
            ptrdiff_t                     n_elements() const { return n_elements_; }

        private:
            std::string           name_;
            ElemType              type_;
            SharedBuffer<idx_t *> elements_;
            ptrdiff_t             n_elements_;
        };

        ptrdiff_t                                   n_nodes() const { return n_nodes_; }
        int                                         spatial_dimension() const { return 3; }
        const SharedBuffer<geom_t *>               &points() const { return points_; }
        const std::vector<std::shared_ptr<Block>> &blocks() const { return blocks_; }
        size_t                                      n_blocks() const { return blocks_.size(); }
        std::shared_ptr<Block>                      block(size_t i) const { return blocks_[i]; }
        std::shared_ptr<Block>                      find_block(const std::string &name) const {
            for (auto &b : blocks_)
                if (b->name() == name) return b;
            return nullptr;
        }
        ptrdiff_t n_elements() const {
            ptrdiff_t n = 0;
            for (auto &b : blocks_) n += b->n_elements();
            return n;
        }

        /* test-harness construction */
        ptrdiff_t                           n_nodes_{0};
        int                                 level_{0}; /* > 0: semi-structured */
        SharedBuffer<geom_t *>              points_;
        std::vector<std::shared_ptr<Block>> blocks_;
    };
}  // namespace smesh
