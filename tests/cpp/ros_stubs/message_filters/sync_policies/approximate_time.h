#pragma once
#include <message_filters/sync_policies/exact_time.h>
namespace message_filters { namespace sync_policies {
template <class M0, class M1, class M2 = NullType, class M3 = NullType> struct ApproximateTime { explicit ApproximateTime(unsigned queue_size); };
} }
