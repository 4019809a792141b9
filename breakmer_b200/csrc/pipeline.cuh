// Batched whole-path pipeline state (declared here, defined in pipeline_impl.cuh).
#pragma once
#include "host_util.cuh"
namespace bk {
struct Pipeline;
}
