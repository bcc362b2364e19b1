#pragma once

#include <faabric/endpoint/FaabricEndpoint.h>

namespace faabric::planner {

// JSON-over-HTTP control API of the planner (reference:
// src/planner/PlannerEndpointHandler.cpp:15-421): body = HttpMessage JSON
class PlannerEndpointHandler final : public faabric::endpoint::HttpRequestHandler
{
  public:
    void onRequest(const faabric::endpoint::HttpRequest& request,
                   faabric::endpoint::HttpResponse& response) override;
};

}
