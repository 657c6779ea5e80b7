// mpt_internal.h — the two things the other translation units of libmpt.so need from a context (not exported).
#pragma once

#include "../../include/mpt.h"

int mpt_internal_device(const mpt_context* ctx);
mpt_status mpt_internal_fail(mpt_context* ctx, mpt_status status, const char* message); // sets mpt_last_error(ctx), returns status
