#pragma once
#include "../sfem_defs.hpp"
