#pragma once
namespace bk { struct Pipeline { int dummy; }; }
namespace {
void pipeline_upload(bk_handle_t, const bk_batch_input*) { fail(BK_ERR_ARG, "not implemented"); }
void pipeline_run(bk_handle_t, const bk_batch_input*, bool, bk_batch_result*) { fail(BK_ERR_ARG, "not implemented"); }
}
