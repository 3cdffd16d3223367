#include <gtsam/stub_core.h>
