// ORACLE — TEST INFRASTRUCTURE ONLY: scene.cpp includes <vk_mem_alloc.h> for the VMA_* enumerators, which the stub gfx/vk.h defines.
#pragma once
#include <gfx/vk.h>
