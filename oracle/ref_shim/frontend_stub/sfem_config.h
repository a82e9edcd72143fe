#pragma once
#include "../sfem_config.h"
