// Oracle build shim (test infrastructure, written for this repo): a no-op stand-in for
// XBot::MatLogger2 so that unmodified reference sources compile without matlogger2.
#pragma once
#include <memory>
#include <string>
namespace XBot {
class MatLogger2 {
public:
    typedef std::shared_ptr<MatLogger2> Ptr;
    template <class T> bool add(const std::string&, const T&) { return true; }
    template <class T> bool create(const std::string&, T, int = 1, int = 0) { return true; }
};
}
