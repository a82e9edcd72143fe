#pragma once
#define SMESH_TRACE_SCOPE(name)
#define SMESH_TRACE_SCOPE_VARIANT(format, num)
