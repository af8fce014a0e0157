// Oracle build shim (test infrastructure): boost::any -> std::any.
#pragma once
#include <any>
namespace boost { using any = std::any; using std::any_cast; using bad_any_cast = std::bad_any_cast; }
