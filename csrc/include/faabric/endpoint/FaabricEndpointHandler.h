#pragma once

#include <faabric/endpoint/FaabricEndpoint.h>

namespace faabric::endpoint {

// Worker-side handler: workers do not take HTTP requests (the planner does),
// everything is rejected (reference: src/endpoint/FaabricEndpointHandler.cpp)
class FaabricEndpointHandler final : public HttpRequestHandler
{
  public:
    void onRequest(const HttpRequest& request, HttpResponse& response) override;
};

}
