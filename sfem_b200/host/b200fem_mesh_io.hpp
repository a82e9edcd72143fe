/* b200fem_mesh_io.hpp -- the reference's on-disk mesh format, host side (C++17, header only).
 *
 * A mesh is a folder: `x.raw y.raw z.raw` (geom_t = fp32, one array per coordinate), `i0.raw .. i<nxe-1>.raw` (idx_t = int32,
 * one array per element corner: the SoA connectivity the operator plans of include/b200fem.h take as they are) and
 * `meta.yaml` naming them (python/sfem/mesh/box_mesh.py:786-806 writes it; the C++ side reads the arrays by pattern,
 * frontend/sfem_Function.cpp:679-727 `i*.raw`, and the coordinates through smesh). Only the flat subset of YAML that file
 * uses is parsed: `key: value` lines and `- name: file` list items under `points:` / `elements:`.
 * tests/golden/ref_mesh/ holds two folders written by the reference's own generator; tests/test_mesh_io.py reads them. */
#ifndef B200FEM_MESH_IO_HPP
#define B200FEM_MESH_IO_HPP

#include <cstdint>
#include <cstdio>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace b200fem {

    struct RawMesh {
        std::string                       element_type; /* "TET4", "HEX8", ... as written in meta.yaml */
        int                               spatial_dimension{3};
        std::vector<std::vector<int32_t>> elements; /* [nxe][ne] */
        std::vector<std::vector<float>>   points;   /* [dim][nn] */
        std::ptrdiff_t n_elements() const { return elements.empty() ? 0 : (std::ptrdiff_t)elements[0].size(); }
        std::ptrdiff_t n_nodes() const { return points.empty() ? 0 : (std::ptrdiff_t)points[0].size(); }
        int            nxe() const { return (int)elements.size(); }
    };

    namespace mesh_io_detail {
        inline std::string trim(const std::string &s) {
            const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
            return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
        }
        template <typename T>
        int read_array(const std::string &path, std::vector<T> &out, std::string *err) {
            std::ifstream f(path, std::ios::binary | std::ios::ate);
            if (!f) {
                *err = "cannot open " + path;
                return 1;
            }
            const std::streamsize bytes = f.tellg();
            if (bytes % (std::streamsize)sizeof(T)) {
                *err = path + ": size is not a multiple of the element size";
                return 1;
            }
            out.resize((size_t)bytes / sizeof(T));
            f.seekg(0);
            f.read(reinterpret_cast<char *>(out.data()), bytes);
            return f ? 0 : (*err = "short read on " + path, 1);
        }
        template <typename T>
        int write_array(const std::string &path, const std::vector<T> &a, std::string *err) {
            std::ofstream f(path, std::ios::binary);
            f.write(reinterpret_cast<const char *>(a.data()), (std::streamsize)(a.size() * sizeof(T)));
            return f ? 0 : (*err = "cannot write " + path, 1);
        }
    }  // namespace mesh_io_detail

    /* returns 0 on success (SFEM_SUCCESS), 1 with *err set otherwise */
    inline int read_mesh(const std::string &folder, RawMesh &m, std::string *err) {
        using namespace mesh_io_detail;
        std::ifstream meta(folder + "/meta.yaml");
        std::map<std::string, std::string>               scalars;
        std::vector<std::pair<std::string, std::string>> pts, els;
        std::string                                       section;
        if (meta) {
            std::string line;
            while (std::getline(meta, line)) {
                const size_t hash = line.find('#');
                if (hash != std::string::npos) line = line.substr(0, hash);
                if (trim(line).empty()) continue;
                const bool  item = trim(line)[0] == '-';
                std::string body = trim(line);
                if (item) body = trim(body.substr(1));
                const size_t colon = body.find(':');
                if (colon == std::string::npos) continue;
                const std::string key = trim(body.substr(0, colon)), val = trim(body.substr(colon + 1));
                if (item) {
                    if (section == "points") pts.emplace_back(key, val);
                    if (section == "elements") els.emplace_back(key, val);
                } else if (val.empty()) {
                    section = key;
                } else {
                    section.clear();
                    scalars[key] = val;
                }
            }
        }
        m = RawMesh();
        if (scalars.count("element_type")) m.element_type = scalars["element_type"];
        if (scalars.count("spatial_dimension")) m.spatial_dimension = std::stoi(scalars["spatial_dimension"]);
        if (pts.empty()) { /* no meta file: the reference's pattern-based reader (x/y/z.raw, i<k>.raw) */
            for (const char *c : {"x", "y", "z"}) pts.emplace_back(c, std::string(c) + ".raw");
            for (int k = 0; k < 64; ++k) {
                const std::string name = "i" + std::to_string(k) + ".raw";
                if (!std::ifstream(folder + "/" + name)) break;
                els.emplace_back("i" + std::to_string(k), name);
            }
        }
        for (auto &p : pts) {
            m.points.emplace_back();
            if (read_array(folder + "/" + p.second, m.points.back(), err)) return 1;
        }
        for (auto &e : els) {
            m.elements.emplace_back();
            if (read_array(folder + "/" + e.second, m.elements.back(), err)) return 1;
        }
        for (auto &p : m.points)
            if (p.size() != m.points[0].size()) return (*err = "coordinate arrays differ in length", 1);
        for (auto &e : m.elements)
            if (e.size() != m.elements[0].size()) return (*err = "connectivity arrays differ in length", 1);
        if (scalars.count("n_elements") && std::stoll(scalars["n_elements"]) != (long long)m.n_elements())
            return (*err = "meta.yaml n_elements does not match the arrays", 1);
        if (scalars.count("n_points") && std::stoll(scalars["n_points"]) != (long long)m.n_nodes())
            return (*err = "meta.yaml n_points does not match the arrays", 1);
        const std::ptrdiff_t nn = m.n_nodes();
        for (auto &e : m.elements)
            for (int32_t v : e)
                if (v < 0 || v >= nn) return (*err = "connectivity refers to a node outside the coordinate arrays", 1);
        if (m.element_type.empty()) m.element_type = m.nxe() == 4 ? "TET4" : (m.nxe() == 8 ? "HEX8" : (m.nxe() == 10 ? "TET10" : "UNKNOWN"));
        return 0;
    }

    inline int write_mesh(const std::string &folder, const RawMesh &m, std::string *err) {
        using namespace mesh_io_detail;
        const char   *axes[3] = {"x", "y", "z"};
        std::ofstream meta(folder + "/meta.yaml");
        if (!meta) return (*err = "cannot write " + folder + "/meta.yaml (the folder must exist)", 1);
        meta << "# SFEM mesh meta file (written by b200fem)\n";
        meta << "elem_num_nodes: " << m.nxe() << "\nelement_type: " << m.element_type << "\nelements:\n";
        for (int k = 0; k < m.nxe(); ++k) meta << "- i" << k << ": i" << k << ".raw\n";
        meta << "n_elements: " << m.n_elements() << "\nn_points: " << m.n_nodes() << "\npoints:\n";
        for (size_t d = 0; d < m.points.size(); ++d) meta << "- " << axes[d] << ": " << axes[d] << ".raw\n";
        meta << "rpath: true\nspatial_dimension: " << m.spatial_dimension << "\n";
        for (size_t d = 0; d < m.points.size(); ++d)
            if (write_array(folder + "/" + axes[d] + ".raw", m.points[d], err)) return 1;
        for (int k = 0; k < m.nxe(); ++k)
            if (write_array(folder + "/i" + std::to_string(k) + ".raw", m.elements[(size_t)k], err)) return 1;
        return 0;
    }

}  // namespace b200fem
#endif
