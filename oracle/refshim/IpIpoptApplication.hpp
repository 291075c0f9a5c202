#include <IpTNLP.hpp>
