// Oracle build shim: intentionally empty.
#pragma once
#include <memory>
