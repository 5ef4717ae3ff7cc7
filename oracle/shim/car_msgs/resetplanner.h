#pragma once
namespace car_msgs { struct resetplanner { struct Request {}; struct Response {}; }; }
