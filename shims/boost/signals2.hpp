// stand-in for boost/signals2.hpp: only what the vendored soth (external/soth-ext) uses; TEST INFRASTRUCTURE (oracle build)
#pragma once
#include <functional>
#include <vector>
namespace boost { using std::function; namespace signals2 {
template <class Sig> class signal;
template <class R, class... A> class signal<R(A...)> {
    std::vector<std::function<R(A...)>> slots_;
public:
    template <class F> void connect(F f) { slots_.emplace_back(f); }
    void operator()(A... a) const { for (auto& s : slots_) s(a...); }
};
} }
