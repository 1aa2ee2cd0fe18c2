// refshim: MVE util/timer.h stand-in (see ../README.md)
#pragma once
#include <chrono>
#include <cstddef>

namespace util {

class WallTimer {
    std::chrono::steady_clock::time_point start;
public:
    WallTimer() { reset(); }
    void reset() { start = std::chrono::steady_clock::now(); }
    std::size_t get_elapsed() const {
        return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - start).count();
    }
    float get_elapsed_sec() const { return get_elapsed() / 1000.0f; }
};
typedef WallTimer ClockTimer;

}  // namespace util
