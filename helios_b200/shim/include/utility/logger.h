// utility/logger.h — the logging macros the engine code on the path uses (reference: include/utility/logger.h);
// the shim writes to stderr.  HELIOS_LOG_FATAL is followed by a throw at every call site, as in the reference
// (e.g. src/engine/gfx/vk.cpp:3320-3374).
#pragma once
#include <cstdio>
#include <string>

#define HELIOS_LOG_INFO(x) std::fprintf(stderr, "[helios][info] %s\n", std::string(x).c_str())
#define HELIOS_LOG_WARNING(x) std::fprintf(stderr, "[helios][warning] %s\n", std::string(x).c_str())
#define HELIOS_LOG_ERROR(x) std::fprintf(stderr, "[helios][error] %s\n", std::string(x).c_str())
#define HELIOS_LOG_FATAL(x) std::fprintf(stderr, "[helios][fatal] %s\n", std::string(x).c_str())
