// Oracle build shim (test infrastructure): dlopen + dlsym + call, the behaviour
// BackEndFactory relies on (reference src/solvers/BackEndFactory.cpp:4-17).
#pragma once
#include <dlfcn.h>
#include <stdexcept>
#include <string>
namespace XBot { namespace Utils {
template <class R, class... Args>
R CallFunction(std::string lib, std::string fn, Args... args)
{
    void* h = dlopen(lib.c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (!h) throw std::runtime_error(std::string("dlopen failed: ") + dlerror());
    typedef R (*fn_t)(Args...);
    fn_t f = (fn_t)dlsym(h, fn.c_str());
    if (!f) throw std::runtime_error(std::string("dlsym failed: ") + fn);
    return f(args...);
}
}}
