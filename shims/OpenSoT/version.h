// Oracle build shim (stands in for the cmake-generated version.h).
#pragma once
#define OPENSOT_VERSION_MAJOR 3
#define OPENSOT_VERSION_MINOR 0
#define OPENSOT_VERSION_PATCH 0
