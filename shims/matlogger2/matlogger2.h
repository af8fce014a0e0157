// Oracle build shim (test infrastructure, written for this repo): a stand-in for XBot::MatLogger2 so that
// unmodified reference sources compile without matlogger2. Nothing is written anywhere; the variable names and a
// checksum of every add() are kept so that tests can compare what two solvers log (tests/cpp/test_dropin.cpp).
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>
namespace XBot {
class MatLogger2 {
public:
    typedef std::shared_ptr<MatLogger2> Ptr;
    struct Entry { std::string name; long size; double checksum; };
    template <class T> bool add(const std::string& name, const T& v)
    {
        entries.push_back(Entry{name, size_of(v, 0), sum_of(v, 0)});
        return true;
    }
    template <class T> bool create(const std::string&, T, int = 1, int = 0) { return true; }
    std::vector<Entry> entries;
private:
    // Eigen types: entries beyond +-1e19 (the reference's clipped infinities, QPOasesBackEnd.cpp:339-356) count as marks
    template <class T> static auto size_of(const T& v, int) -> decltype(v.size(), long()) { return (long)v.size(); }
    template <class T> static long size_of(const T&, long) { return 1; }
    template <class T> static auto sum_of(const T& v, int) -> decltype(v.sum(), double())
    {
        double s = 0.0;
        for (long j = 0; j < (long)v.cols(); ++j)
            for (long i = 0; i < (long)v.rows(); ++i) {
                const double x = (double)v(i, j);
                s += x > 1e19 ? 7.0 : (x < -1e19 ? -11.0 : x);
            }
        return s;
    }
    template <class T> static double sum_of(const T&, long) { return 0.0; }
};
}
