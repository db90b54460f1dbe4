/* TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for the CUDA-samples header that the
 * reference's kernel.cu includes (kernel.cu:4) and uses at kernel.cu:298-301,482-484.
 * Written from the six call signatures used there; not derived from the CUDA samples source. */
#pragma once
#include <chrono>

struct StopWatchInterface
{
    std::chrono::steady_clock::time_point t0;
    double acc_ms = 0.0;
    bool running = false;
};

static inline bool sdkCreateTimer( StopWatchInterface** t ) { *t = new StopWatchInterface(); return true; }
static inline bool sdkDeleteTimer( StopWatchInterface** t ) { delete *t; *t = nullptr; return true; }
static inline bool sdkResetTimer( StopWatchInterface** t ) { ( *t )->acc_ms = 0.0; ( *t )->running = false; return true; }
static inline bool sdkStartTimer( StopWatchInterface** t )
{
    ( *t )->t0 = std::chrono::steady_clock::now();
    ( *t )->running = true;
    return true;
}
static inline bool sdkStopTimer( StopWatchInterface** t )
{
    if( ( *t )->running )
    {
        ( *t )->acc_ms += std::chrono::duration< double, std::milli >( std::chrono::steady_clock::now() - ( *t )->t0 ).count();
        ( *t )->running = false;
    }
    return true;
}
static inline float sdkGetTimerValue( StopWatchInterface** t ) { return ( float )( *t )->acc_ms; }
