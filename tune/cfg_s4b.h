#define YF_CFGRES1_W WirbCfg<4,8,4,1,8,8,6,10,true,true>
#define YF_CFGRES2_W WirbCfg<8,32,8,1,4,2,6,12,true,false>
#define YF_CFGDOWN2_W WirbCfg<8,32,8,2,4,2,8,16,false,false>
