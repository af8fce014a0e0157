// Oracle build shim (test infrastructure): silent XBot::Logger with the call shapes the
// reference uses (printf-style statics and stream-style error() << ... << endl()).
#pragma once
#include <cstdarg>
#include <cstdio>
#include <iostream>
#include <ostream>
#include <sstream>
#include <string>
#include <matlogger2/matlogger2.h>
namespace XBot {
class Logger {
public:
    enum class Severity { LOW = 0, MID = 1, HIGH = 2, FATAL = 3, DEBUG = -1 };
    struct Sink {
        template <class T> Sink& operator<<(const T&) { return *this; }
        Sink& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
    };
    struct Endl {};
    static Sink& sink() { static Sink s; return s; }
    static Sink& info() { return sink(); }
    static Sink& success() { return sink(); }
    static Sink& warning() { return sink(); }
    static Sink& error() { return sink(); }
    static Sink& info(Severity) { return sink(); }
    static Sink& error(Severity) { return sink(); }
    static Endl endl() { return Endl(); }
    static void info(const char*, ...) {}
    static void success(const char*, ...) {}
    static void warning(const char*, ...) {}
    static void error(const char* fmt, ...) {
        if (!verbose()) return;
        va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
    }
    static void info(Severity, const char*, ...) {}
    static void error(Severity, const char*, ...) {}
    static void SetVerbosityLevel(Severity) {}
    static bool& verbose() { static bool v = false; return v; }
};
}
