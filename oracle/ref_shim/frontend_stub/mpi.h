#pragma once
#include "../mpi.h"
