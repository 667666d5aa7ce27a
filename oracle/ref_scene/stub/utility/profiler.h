// ORACLE — TEST INFRASTRUCTURE ONLY: the reference's profiler (src/engine/utility/profiler.cpp) is Windows + Vulkan timestamp queries.
#pragma once
#define HELIOS_SCOPED_SAMPLE(name)
