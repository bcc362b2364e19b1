// The reference re-exports boost::asio / boost::beast here
// (include/faabric/util/asio.h).  The HTTP endpoint of this tree is
// self-contained (endpoint/FaabricEndpoint.h): the request / response types
// keep the reference's alias names.
#pragma once

#include <faabric/endpoint/FaabricEndpoint.h>

namespace faabric::util {

using BeastHttpRequest = faabric::endpoint::HttpRequest;
using BeastHttpResponse = faabric::endpoint::HttpResponse;

}
