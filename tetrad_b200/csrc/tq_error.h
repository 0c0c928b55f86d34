// Last-error text of the C ABI (tq_last_error), shared by the translation units of libtetrad_b200.so.
#pragma once

namespace tq {

// printf-style: stores the message in the calling thread's error slot and returns `code`.
int set_error(int code, const char *fmt, ...);

}  // namespace tq
