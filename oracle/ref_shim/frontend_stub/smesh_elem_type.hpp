#pragma once
#include "../smesh_elem_type.hpp"
