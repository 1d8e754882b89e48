/*
 * TEST INFRASTRUCTURE ONLY (oracle). No-op stand-in for src-core/logger.h so that the
 * reference's DSP/FEC translation units compile without the logger/sinks implementation.
 */
#pragma once
#include <memory>
#include <string>
namespace slog
{
    class Logger
    {
    public:
        template <typename... A> void trace(A...) {}
        template <typename... A> void debug(A...) {}
        template <typename... A> void info(A...) {}
        template <typename... A> void warn(A...) {}
        template <typename... A> void error(A...) {}
        template <typename... A> void critical(A...) {}
    };
}
extern std::shared_ptr<slog::Logger> logger;
